#!/usr/bin/env python
"""Generates the end-to-end golden fixtures under tests/golden/ by running the REAL reference
(oracle/_ref/winnowmap = /root/reference + the documented rep_len=0 init, built by oracle/build_ref.sh)
on small deterministic synthetic inputs.  Inputs are regenerated from seeds at test time (tools/gen_data.py);
the manifest stores their md5 so that generator drift is detected.  Run only where /root/reference exists."""
import gzip
import hashlib
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import gen_data  # noqa: E402
import numpy as np  # noqa: E402

CASES = {
    # name: (ref_len, contigs, tandem, ref_seed, n_reads, n50, err, read_seed, min_len, preset, use_W, extra)
    "ont_small": dict(ref_len=300000, contigs=2, tandem=False, ref_seed=1011, n_reads=60, n50=9000, err=0.05, read_seed=2011, min_len=500,
                      preset="map-ont", use_W=False, k=15),
    "ont_tandem": dict(ref_len=400000, contigs=2, tandem=True, ref_seed=1005, n_reads=60, n50=12000, err=0.05, read_seed=2005, min_len=1000,
                       preset="map-ont", use_W=True, k=15),
    # w_distinct: the `distinct=` argument of the meryl rule; 0.99 here so that the list holds the k-mers of the planted array
    "hifi_small": dict(ref_len=300000, contigs=1, tandem=True, ref_seed=1013, n_reads=40, n50=12000, err=0.005, read_seed=2013, min_len=1000,
                       preset="map-pb", use_W=True, k=15, w_distinct=0.99),
    # reads drawn from a donor genome carrying deletions / insertions / inversions: Z-drop splits, long joins, inversion rescue
    "ont_sv": dict(ref_len=400000, contigs=2, tandem=False, ref_seed=1021, n_reads=60, n50=14000, err=0.04, read_seed=2021, min_len=2000,
                   preset="map-ont", use_W=False, k=15, sv=True),
    # a 300-bp family present > 5000 times: minimizers above mid_occ are skipped and rl:i: becomes non-zero
    "ont_highocc": dict(ref_len=2600000, contigs=1, tandem=False, ref_seed=1022, n_reads=40, n50=9000, err=0.05, read_seed=2022, min_len=1500,
                        preset="map-ont", use_W=True, k=15, highocc=True),
    # -O4 -E2: one gap pair (q == q2, e == e2), every DP call is ksw_extz2_sse (src/align.c:328-331)
    # (the structural-variant reads: long gaps are where one gap pair and two differ)
    "ont_single_gap": dict(ref_len=400000, contigs=2, tandem=False, ref_seed=1021, n_reads=60, n50=14000, err=0.04, read_seed=2021, min_len=2000,
                           preset="map-ont", use_W=False, k=15, sv=True, extra=["-O4", "-E2"], gap=(4, 2, 4, 2)),
    "asm20_small": dict(ref_len=300000, contigs=1, tandem=True, ref_seed=1014, n_reads=12, n50=40000, err=0.02, read_seed=2014, min_len=5000,
                        preset="asm20", use_W=True, k=19, w_distinct=0.99),
}


def md5(path):
    return hashlib.md5(open(path, "rb").read()).hexdigest()


def make_inputs(name, outdir):
    c = CASES[name]
    os.makedirs(outdir, exist_ok=True)
    ref = os.path.join(outdir, name + ".ref.fa")
    reads = os.path.join(outdir, name + ".reads.fa")
    wfile = os.path.join(outdir, name + ".rep.txt")
    rng = np.random.default_rng(c["ref_seed"])
    contigs = gen_data.make_ref(rng, c["ref_len"], c["contigs"], c["tandem"])
    if c.get("highocc"):
        name, seq = contigs[0]
        seq = seq.copy()
        unit = gen_data.random_seq(rng, 300)
        pos = 20000
        for _ in range(6200):  # interspersed copies, 2 % divergence each, ~100 bp apart
            cp = gen_data.mutate(rng, unit, 0.02, (1.0, 0.0, 0.0))
            if pos + 300 >= len(seq) - 20000:
                break
            seq[pos:pos + 300] = cp
            pos += 300 + int(rng.integers(60, 140))
        contigs = [(name, seq)]
    gen_data.write_fasta(ref, contigs)
    rng = np.random.default_rng(c["read_seed"])
    donor = contigs
    if c.get("sv"):
        donor = []
        for name, seq in contigs:
            parts, p = [], 0
            while p < len(seq):
                step = int(rng.integers(6000, 16000))
                seg = seq[p:p + step]
                kind = int(rng.integers(0, 4))
                if kind == 0 and len(seg) > 6000:      # deletion
                    d = int(rng.integers(300, 3000))
                    seg = np.concatenate([seg[:2000], seg[2000 + d:]])
                elif kind == 1:                         # insertion of novel sequence
                    seg = np.concatenate([seg[:2500], gen_data.random_seq(rng, int(rng.integers(300, 2500))), seg[2500:]])
                elif kind == 2 and len(seg) > 6000:    # inversion
                    L = int(rng.integers(800, 3000))
                    seg = np.concatenate([seg[:2000], gen_data.COMP[seg[2000:2000 + L][::-1]], seg[2000 + L:]])
                parts.append(seg)
                p += step
            donor.append((name, np.concatenate(parts)))
    recs = gen_data.make_reads(rng, donor, c["n_reads"], c["n50"], c["err"], min_len=c["min_len"])
    gen_data.write_fasta(reads, recs)
    if c["use_W"]:
        gen_data.write_top_kmers(wfile, contigs, c["k"], c.get("w_distinct", 0.9998))
    else:
        wfile = None
    return ref, reads, wfile


SAM_CASES = ["ont_small", "ont_sv"]
# key -> (case, reference options); flags for the library: MM_F_OUT_CS 0x40, MM_F_OUT_CS_LONG 0x800, MM_F_OUT_MD 0x1000000
TAG_CASES = {
    "paf_cs": ("ont_small", ["-c", "--cs"]),
    "paf_cs_long": ("ont_small", ["-c", "--cs=long"]),
    "sam_md": ("ont_sv", ["-a", "--MD"]),
    "paf_eqx": ("ont_sv", ["-c", "--eqx"]),                      # MM_F_EQX 0x4000000
    "sam_softclip": ("ont_sv", ["-a", "-Y"]),                    # MM_F_SOFTCLIP 0x80000
    "sam_no2nd_hitonly": ("ont_highocc", ["-a", "--secondary=no", "--sam-hit-only"]),  # 0x4000 | 0x40000000
    "paf_no_hit": ("ont_highocc", ["-c", "--paf-no-hit"]),       # MM_F_PAF_NO_HIT 0x8000000
    "sam_fastq_comment": ("ont_small", ["-a", "-y"], "fastq"),   # gzipped FASTQ with comments: QUAL column, MM_F_COPY_COMMENT 0x2000000
    "paf_edge": ("ont_small", ["-c"], "edge"),                   # empty / tiny / N-rich / lower-case / chimeric / unmappable reads
    "sam_edge": ("ont_small", ["-a"], "edge"),
}


def read_fasta(path):
    recs, name, seq = [], None, []
    for ln in open(path):
        ln = ln.rstrip("\n")
        if ln.startswith(">"):
            if name is not None:
                recs.append((name, "".join(seq)))
            name, seq = ln[1:], []
        else:
            seq.append(ln)
    if name is not None:
        recs.append((name, "".join(seq)))
    return recs


def edge_reads_of(reads_fa, out):
    """Awkward inputs derived from the case's reads: empty, shorter than k, around k, short, lower case, N runs and IUPAC
    codes, an unmappable random read, a chimera of two reads (second half reverse-complemented), a long read with a big
    N block, and two untouched reads for reference."""
    recs = read_fasta(reads_fa)
    rng = np.random.default_rng(77)
    comp = str.maketrans("ACGTacgt", "TGCAtgca")
    a, b, c = recs[0][1], recs[1][1], max(recs, key=lambda r: len(r[1]))[1]
    out_recs = [
        ("e_empty", ""), ("e_len10", a[100:110]), ("e_len14", a[200:214]), ("e_len15", a[300:315]), ("e_len30", a[400:430]),
        ("e_len200", a[500:700]), ("e_len999", b[100:1099]), ("e_lower", b.lower()),
        ("e_nrun", a[:1500] + "N" * 300 + a[1800:]), ("e_iupac", b[:800] + "RYKMSWN" * 20 + b[940:]),
        ("e_random", "".join("ACGT"[i] for i in rng.integers(0, 4, 5000))),
        ("e_chimera", a[:len(a) // 2] + b[:len(b) // 2].translate(comp)[::-1]),
        ("e_bigN", c[:4000] + "N" * 3000 + c[7000:]),
        ("e_plain0", a), ("e_plain1", b),
    ]
    with open(out, "w") as f:
        for nm, sq in out_recs:
            f.write(f">{nm}\n")
            for i in range(0, len(sq), 80):
                f.write(sq[i:i + 80] + "\n")
    return out



def fastq_gz_of(reads_fa, out):
    """The reads as gzipped FASTQ with a comment and a deterministic quality string (input for the QUAL / -y tests)."""
    recs, name, seq = [], None, []
    for ln in open(reads_fa):
        ln = ln.rstrip("\n")
        if ln.startswith(">"):
            if name is not None:
                recs.append((name, "".join(seq)))
            name, seq = ln[1:], []
        else:
            seq.append(ln)
    if name is not None:
        recs.append((name, "".join(seq)))
    with gzip.GzipFile(out, "wb", mtime=0) as f:
        for i, (nm, sq) in enumerate(recs):
            q = "".join(chr(33 + (7 * j + 3 * i) % 41) for j in range(len(sq)))
            f.write(f"@{nm} RG:Z:grp{i % 3}\tXX:i:{i}\n{sq}\n+\n{q}\n".encode())
    return out



def sam_without_pg(sam):
    """The @PG line records the command line (temporary paths): everything else must match byte for byte."""
    return b"".join(ln for ln in sam.splitlines(keepends=True) if not ln.startswith(b"@PG"))


def sam_strip_seq(sam):
    """SEQ and QUAL replaced by their lengths: a small committed fixture to locate a difference; the full text is
    pinned by its md5 in the manifest."""
    out = []
    for ln in sam.splitlines(keepends=True):
        if ln.startswith(b"@"):
            out.append(ln)
            continue
        f = ln.rstrip(b"\n").split(b"\t")
        f[9] = b"len=%d" % len(f[9]) if f[9] != b"*" else b"*"
        f[10] = b"len=%d" % len(f[10]) if f[10] != b"*" else b"*"
        out.append(b"\t".join(f) + b"\n")
    return b"".join(out)


def main():
    refbin = os.path.join(ROOT, "oracle", "_ref", "winnowmap")
    if not os.path.exists(refbin):
        subprocess.check_call([os.path.join(ROOT, "oracle", "build_ref.sh")])
    gdir = os.path.join(ROOT, "tests", "golden")
    os.makedirs(gdir, exist_ok=True)
    tmp = "/tmp/wm_golden"
    manifest = {}
    for name, c in CASES.items():
        ref, reads, wfile = make_inputs(name, tmp)
        cmd = [refbin, "-t", "4", "-c", "-x", c["preset"]] + c.get("extra", [])
        if wfile:
            cmd += ["-W", wfile]
        cmd += [ref, reads]
        out = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, check=True).stdout
        with gzip.GzipFile(os.path.join(gdir, name + ".paf.gz"), "wb", mtime=0) as f:
            f.write(out)
        manifest[name] = dict(params=c, ref_md5=md5(ref), reads_md5=md5(reads), w_md5=md5(wfile) if wfile else None,
                              cmd=" ".join(["winnowmap"] + cmd[1:]), n_lines=out.count(b"\n"), paf_md5=hashlib.md5(out).hexdigest())
        print(name, manifest[name]["n_lines"], "lines", len(out), "bytes")
    for name in SAM_CASES:  # the same inputs with -a: SAM records (flags, clipping, SEQ/QUAL, SA:Z) and @SQ header
        c = CASES[name]
        ref, reads, wfile = make_inputs(name, tmp)
        cmd = [refbin, "-t", "4", "-a", "-x", c["preset"]]
        if wfile:
            cmd += ["-W", wfile]
        cmd += [ref, reads]
        out = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, check=True).stdout
        body = sam_without_pg(out)
        with gzip.GzipFile(os.path.join(gdir, name + ".sam.stripped.gz"), "wb", mtime=0) as f:
            f.write(sam_strip_seq(body))
        manifest[name]["sam_md5"] = hashlib.md5(body).hexdigest()
        manifest[name]["sam_lines"] = body.count(b"\n")
        print(name, "SAM", manifest[name]["sam_lines"], "lines", len(body), "bytes")
    for key, case in TAG_CASES.items():  # output options: --cs, --cs=long, --MD, --eqx, -Y, ...
        name, args = case[0], case[1]
        c = CASES[name]
        ref, reads, wfile = make_inputs(name, tmp)
        if len(case) > 2 and case[2] == "fastq":
            reads = fastq_gz_of(reads, reads + ".fq.gz")
        if len(case) > 2 and case[2] == "edge":
            reads = edge_reads_of(reads, reads + ".edge.fa")
        cmd = [refbin, "-t", "4", "-x", c["preset"]] + args
        if wfile:
            cmd += ["-W", wfile]
        cmd += [ref, reads]
        out = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, check=True).stdout
        body = sam_without_pg(out)
        manifest[name].setdefault("tag_md5", {})[key] = hashlib.md5(body).hexdigest()
        if key == "paf_cs":
            with gzip.GzipFile(os.path.join(gdir, name + ".cs.paf.gz"), "wb", mtime=0) as f:
                f.write(body)
        print(name, key, body.count(b"\n"), "lines", len(body), "bytes")
    json.dump(manifest, open(os.path.join(gdir, "manifest.json"), "w"), indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
