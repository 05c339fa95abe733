import sys, numpy as np
a = np.loadtxt(sys.argv[1], delimiter=",", dtype=np.int64)
a = a[a[:, 0] > 0]
t0, t1, t2, t3, t4, sm, wa, wb = a.T
print("CTAs", len(a))
for name, d in (("prologue (start -> after TMEM alloc / tables)", t1 - t0), ("first halo landed", t2 - t1), ("main loop (first halo -> accumulator complete)", t3 - t2), ("epilogue", t4 - t3), ("CTA lifetime", t4 - t0), ("  MMA thread waiting for later halos", wa), ("  MMA thread waiting for weight stages", wb)):
    print(f"{name:50s} mean {d.mean():9.0f} cyc  median {np.median(d):9.0f}  p90 {np.percentile(d,90):9.0f}")
