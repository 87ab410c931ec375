"""Developer tool: SUO_TRACE=<prefix> python bench.py ... writes <prefix>.<ctx>.csv (globaltimer at the start / end of CTA 0 of every
persistent conv launch).  This script sums, over the steady-state steps, the time between the end of one conv kernel and the start of
the next one (launch gap + whatever non-conv kernel ran in between) and the kernel durations.  usage: python tools/trace_gaps.py file.csv"""
import sys
import numpy as np
d = np.loadtxt(sys.argv[1], delimiter=",", skiprows=1)
d = d[d[:, 2] > 0]
t0, t1 = d[:, 1], d[:, 2]
per_step = 187
n = (len(d) // per_step) * per_step
steps = n // per_step
print("launches", len(d), "steps", steps)
dur = (t1 - t0)[:n].reshape(steps, per_step) * 1e-3
gap = (t0[1:] - t1[:-1])
gap = np.concatenate([gap, [0]])[:n].reshape(steps, per_step) * 1e-3
s = slice(5, steps - 1)
print("per step (us): kernel time %.0f, gaps %.0f (median over steps)" % (np.median(dur[s].sum(1)), np.median(gap[s][:, :-1].sum(1))))
g = np.median(gap[s], axis=0)[:-1]
print("gap per launch (us): median %.2f, p10 %.2f, p90 %.2f, max %.1f" % (np.median(g), np.percentile(g, 10), np.percentile(g, 90), g.max()))
print("gaps > 10 us (index: us):", {int(i): round(float(g[i]), 1) for i in np.where(g > 10)[0]})
k = np.median(dur[s], axis=0)
print("kernels < 20 us: %d launches, %.0f us in total; their following gaps %.0f us" % ((k < 20).sum(), k[k < 20].sum(), g[(k < 20)[:-1]].sum()))
