"""Timeline of CTA pair 0 inside one K1 launch (clock64 stamps, leader SM).
usage: match_timeline.py n [rows_to_print]"""
import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import sfm_mvs_b200 as sfm
from sfm_mvs_b200 import synth
from sfm_mvs_b200._lib import lib, check
n = int(sys.argv[1]) if len(sys.argv) > 1 else 5000
show = int(sys.argv[2]) if len(sys.argv) > 2 else 40
ctx = sfm.Context(0)
q = ctx.descriptors(synth.sift_like_descriptors(n, seed=0))
t = ctx.descriptors(synth.sift_like_descriptors(n, seed=1))
buf = np.zeros((4096, 8), np.int64)
info = np.zeros(5, np.int32)
for _ in range(3):
    check(lib.sfm_debug_match_tc_timeline(ctx._h, q._h, t._h, buf.ctypes.data, buf.shape[0], info.ctypes.data))
qt, nsplit, n_items, n_pairs, jobs = (int(x) for x in info)
print(f"n {n}: QT {qt} nsplit {nsplit} items {n_items} pairs {n_pairs} jobs(pair0) {jobs}")
s = buf[:jobs].astype(np.float64)
t0 = s[s > 0].min()
names = ["Bown", "Bpeer", "ready", "accfree", "issued", "accfull", "released", "folded"]
print("job  " + " ".join(f"{x:>9s}" for x in names) + " | wait_B wait_acc mma->full full->rel fold")
for j in range(min(jobs, show)):
    r = s[j] - t0
    print(f"{j:3d}  " + " ".join(f"{x:9.0f}" for x in r) +
          f" | {r[1]-r[0]:6.0f} {r[3]-r[2]:7.0f} {r[5]-r[4]:8.0f} {r[6]-r[5]:8.0f} {r[7]-r[6]:5.0f}")
if qt == 2:
    lat = [s[2 * k, 0] - s[2 * k + 1, 0] for k in range(jobs // 2) if s[2 * k + 1, 0] > 0]
    print("train-tile load: issue -> landed (cycles):", " ".join(f"{x:.0f}" for x in lat[:24]))
if jobs > 2:
    d = np.diff(s[:jobs, 4])
    print("issue-to-issue cycles: median %.0f mean %.0f min %.0f max %.0f; total %.0f cycles for %d jobs" %
          (np.median(d), d.mean(), d.min(), d.max(), s[jobs - 1, 7] - t0, jobs))
