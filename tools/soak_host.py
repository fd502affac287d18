#!/usr/bin/env python
"""Soak test of the host orchestration (wave scheduler, alignment state machines, glue, writers): random small data sets,
the reference binary against the product's host sources on the oracle backend (tests/hostsim), byte comparison.
CPU only; needs oracle/_ref/winnowmap.  Usage: soak_host.py [n_datasets] [first_seed]"""
import ctypes as C
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import gen_data  # noqa: E402
import make_golden  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 10
    first = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    subprocess.check_call([os.path.join(ROOT, "tests", "hostsim", "build.sh")], stderr=subprocess.DEVNULL)
    L = C.CDLL(os.path.join(ROOT, "tests", "hostsim", "libwm_hostsim.so"))
    L.wmt_map_file_flags.argtypes = [C.c_char_p] * 5 + [C.c_int, C.c_int, C.c_int64]
    refbin = os.path.join(ROOT, "oracle", "_ref", "winnowmap")
    out = "/tmp/wm_soak"
    os.makedirs(out, exist_ok=True)
    bad = 0
    for seed in range(first, first + n):
        rng = np.random.default_rng(50000 + seed)
        preset = ["map-ont", "map-ont", "map-pb", "asm20"][seed % 4]
        case = dict(ref_len=int(rng.integers(150000, 600000)), contigs=int(rng.integers(1, 4)), tandem=bool(seed % 3 == 0), ref_seed=60000 + seed,
                    n_reads=int(rng.integers(20, 60)) if preset != "asm20" else 8, n50=int(rng.integers(3000, 16000)) if preset != "asm20" else 30000,
                    err=0.05 if preset == "map-ont" else 0.005 if preset == "map-pb" else 0.02, read_seed=70000 + seed,
                    min_len=int(rng.integers(200, 2000)), preset=preset, use_W=bool(seed % 2), k=19 if preset == "asm20" else 15, sv=bool(seed % 5 == 1))
        name = f"soak{seed}"
        make_golden.CASES[name] = case
        ref, reads, wfile = make_golden.make_inputs(name, out)
        sam = seed % 2
        cmd = [refbin, "-t", "4", "-a" if sam else "-c", "-x", preset] + (["-W", wfile] if wfile else []) + [ref, reads]
        exp = make_golden.sam_without_pg(subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, check=True).stdout)
        o = os.path.join(out, name + ".out")
        rc = L.wmt_map_file_flags(ref.encode(), wfile.encode() if wfile else None, preset.encode(), reads.encode(), o.encode(), 8, sam, 0)
        got = make_golden.sam_without_pg(open(o, "rb").read()) if rc == 0 else b""
        ok = rc == 0 and got == exp
        bad += not ok
        print(f"seed {seed} {preset} {'SAM' if sam else 'PAF'} tandem={case['tandem']} sv={case['sv']} W={case['use_W']} lines={exp.count(10)} -> {'identical' if ok else 'DIFFERENT rc=%d' % rc}", flush=True)
    print("mismatching data sets:", bad)
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
