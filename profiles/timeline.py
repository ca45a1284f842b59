"""Per-tile phase timeline of k_fused_v4 (debug instantiation, OTMB_V4_TIMELINE=<file>).
    OTMB_V4_TIMELINE=gpurun_out/timeline.bin python profiles/prof_driver.py fused 3 && python profiles/timeline.py gpurun_out/timeline.bin
Stamps per tile (64 int64): 0 globaltimer at block start, 1 SM clock at block start, 6 SM id; scan warp: 2 passed barrier 1
(every column warp has published its counts), 3 aggregate published, 4 look-back done, 5 look-back rounds; column warp w:
8+4w counts published, 9+4w reached its release barrier, 10+4w passed it, 11+4w end of the tile.  SM clock = 1.965 GHz."""
import sys
import numpy as np

a = np.fromfile(sys.argv[1], dtype=np.int64).reshape(-1, 64)
nt = a.shape[0]
NW = 11
us = lambda cyc: cyc / 1965.0
t0 = a[:, 1]
w = a[:, 8:8 + 4 * NW].reshape(nt, NW, 4)
full = (w[:, :, 3] > 0).all(axis=1) & (a[:, 4] > 0)      # tiles whose every warp held columns
a, t0, w = a[full], t0[full], w[full]
rel = lambda x: us(x - t0[:, None]) if x.ndim == 2 else us(x - t0)
q = lambda x: "p10 %.2f  p50 %.2f  p90 %.2f  p99 %.2f  mean %.2f" % (*np.percentile(x, [10, 50, 90, 99]), x.mean())
print(f"{full.sum()} of {nt} tiles with all {NW} column warps busy; times in us from block start")
print("counts published (per warp)      ", q(rel(w[:, :, 0]).ravel()))
print("  slowest warp of the tile       ", q(rel(w[:, :, 0]).max(axis=1)))
print("scan warp passed barrier 1       ", q(rel(a[:, 2])))
print("aggregate published              ", q(rel(a[:, 3])))
print("look-back done                   ", q(rel(a[:, 4])))
print("  look-back duration             ", q(us(a[:, 4] - a[:, 3])))
print("  look-back rounds               ", q(a[:, 5].astype(float)))
print("warp reaches its release barrier ", q(rel(w[:, :, 1]).ravel()))
print("  wait at the barrier (per warp) ", q(us(w[:, :, 2] - w[:, :, 1]).ravel()))
print("  warps that wait > 0.2 us       ", "%.1f %%" % (100 * (us(w[:, :, 2] - w[:, :, 1]) > 0.2).mean()))
print("end of the tile (per warp)       ", q(rel(w[:, :, 3]).ravel()))
print("  last warp of the tile          ", q(rel(w[:, :, 3]).max(axis=1)))
print("pre-barrier work (publish->reach)", q(us(w[:, :, 1] - w[:, :, 0]).ravel()))
print("post-barrier work (pass->end)    ", q(us(w[:, :, 3] - w[:, :, 2]).ravel()))
g = a[:, 0] - a[:, 0].min()
print("block starts: first %.1f us, last %.1f us after the first block (globaltimer)" % (g.min() / 1e3, g.max() / 1e3))

# ---- slot turn-around: on each SM two blocks are resident; a new block starts when one of them has retired
sm = a[:, 6]
start = a[:, 1]
end = w[:, :, 3].max(axis=1)
gaps, idle = [], []
for s_ in np.unique(sm):
    sel = np.flatnonzero(sm == s_)
    o = sel[np.argsort(start[sel])]
    ends = []                                    # end clocks of the blocks currently resident on this SM
    for i in o:
        if len(ends) >= 2:
            e = min(ends)
            ends.remove(e)
            gaps.append(us(start[i] - e))
        ends.append(end[i])
    span = end[o].max() - start[o].min()
    idle.append(1.0 - (end[o] - start[o]).sum() / (2.0 * span))
gaps = np.array(gaps)
print("slot turn-around (last warp's end -> next block's first instruction)", q(gaps))
print("resident-slot idle fraction per SM: mean %.1f %%" % (100 * np.mean(idle)))
