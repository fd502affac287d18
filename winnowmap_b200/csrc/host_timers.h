// Tiny named wall-clock accumulators for the orchestration thread (bench / tuning aid; dumped by wm_dump_timers()).
#pragma once
#include <chrono>
#include <stdio.h>
#include <string.h>
#include <mutex>

namespace wmh {
struct Timers {
	enum { MAXT = 48 };
	const char *name[MAXT]; double sec[MAXT]; long cnt[MAXT]; int n;
	Timers() : n(0) {}
	static double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
	std::mutex mu;
	void add(const char *nm, double dt) {
		std::lock_guard<std::mutex> lk(mu);
		for (int i = 0; i < n; ++i) if (name[i] == nm || strcmp(name[i], nm) == 0) { sec[i] += dt; ++cnt[i]; return; }
		if (n < MAXT) { name[n] = nm; sec[n] = dt; cnt[n] = 1; ++n; }
	}
	void dump(FILE *fp) { for (int i = 0; i < n; ++i) fprintf(fp, "[timer] %-28s %9.3f ms  n=%ld\n", name[i], sec[i] * 1e3, cnt[i]); }
	void reset() { n = 0; }
};
extern Timers g_timers;
struct ScopedTimer {
	const char *nm; double t0;
	ScopedTimer(const char *n_) : nm(n_), t0(Timers::now()) {}
	~ScopedTimer() { g_timers.add(nm, Timers::now() - t0); }
};
}
#define WM_TIMED(name) wmh::ScopedTimer wm_scoped_timer_##__LINE__(name)
