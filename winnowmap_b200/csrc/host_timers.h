// Tiny named wall-clock accumulators for the orchestration thread (bench / tuning aid; dumped by wm_dump_timers()).
#pragma once
#include <chrono>
#include <time.h>
#include <stdio.h>
#include <string.h>
#include <stdlib.h>
#include <mutex>
#include <thread>
#include <vector>

namespace wmh {
struct Timers {
	enum { MAXT = 64 };
	const char *name[MAXT]; double sec[MAXT], cpu[MAXT]; long cnt[MAXT]; int n;
	Timers() : n(0) {}
	static double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
	std::mutex mu;
	// CPU time of the whole process (all threads): the cost of a phase when a single lane is running (WM_LANES=1)
	static double cpu_now() { struct timespec ts; clock_gettime(CLOCK_PROCESS_CPUTIME_ID, &ts); return ts.tv_sec + 1e-9 * ts.tv_nsec; }
	// WM_TIMELINE=<file>: every interval is also kept as (thread, name, end, length) and written by dump() -- the lanes' phases on
	// one time axis (tuning aid)
	struct Ev { size_t tid; const char *nm; double t1, dt; };
	std::vector<Ev> log;
	void add(const char *nm, double dt, double dcpu = 0.0) {
		static const bool tl = getenv("WM_TIMELINE") != 0;
		std::lock_guard<std::mutex> lk(mu);
		if (tl) { Ev e; e.tid = std::hash<std::thread::id>()(std::this_thread::get_id()); e.nm = nm, e.t1 = now(), e.dt = dt; log.push_back(e); }
		for (int i = 0; i < n; ++i) if (name[i] == nm || strcmp(name[i], nm) == 0) { sec[i] += dt; cpu[i] += dcpu; ++cnt[i]; return; }
		if (n < MAXT) { name[n] = nm; sec[n] = dt; cpu[n] = dcpu; cnt[n] = 1; ++n; }
	}
	void dump(FILE *fp) {
		if (const char *fn = getenv("WM_TIMELINE")) {
			if (FILE *f = fopen(fn, "a")) {
				for (const Ev &e : log) fprintf(f, "%zu\t%s\t%.6f\t%.6f\n", e.tid, e.nm, e.t1 - e.dt, e.t1);
				fprintf(f, "#dump\n");
				fclose(f);
			}
			log.clear();
		}
		for (int i = 0; i < n; ++i) fprintf(fp, "[timer] %-28s %9.3f ms  cpu %9.3f ms  n=%ld\n", name[i], sec[i] * 1e3, cpu[i] * 1e3, cnt[i]);
	}
	void reset() { n = 0; }
};
extern Timers g_timers;
struct ScopedTimer {
	const char *nm; double t0, c0;
	ScopedTimer(const char *n_) : nm(n_), t0(Timers::now()), c0(Timers::cpu_now()) {}
	~ScopedTimer() { g_timers.add(nm, Timers::now() - t0, Timers::cpu_now() - c0); }
};
}
#define WM_TIMED(name) wmh::ScopedTimer wm_scoped_timer_##__LINE__(name)
