"""DEVELOPER ANALYSIS AID: passes through `pow` per warp, today's control flow vs a per-lane state
machine (every pass serves each lane's next power, whatever its stage).  Uses libemu_stage.so
(g++ -O1 -ffp-contract=off -shared -fPIC emu_stage.cpp -o libemu_stage.so).
python scripts/host_emu/stage_analysis.py [n_cells<=8192] [n_years]"""
import ctypes as C, os, sys
import numpy as np
HERE = os.path.dirname(os.path.abspath(__file__)); ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from pyhelp_b200 import grid as G, _native as N
from support import synth
n = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
ny = int(sys.argv[2]) if len(sys.argv) > 2 else 2
kind = sys.argv[3] if len(sys.argv) > 3 else "natural3"
nn = max(1, n // 100)
res = {"natural3": synth.grid_natural_3layer, "mixed": synth.grid_mixed}[kind](n, seed=1)
g, ex = res if isinstance(res, tuple) else (res, None)
wx = synth.synth_weather(nn, 2001, ny, seed=5)
node = synth.assign_nodes(n, nn)
b = G.batch_from_grid(g, wx["years_of_day"], wx["precip"], wx["airtemp"], wx["solrad"], node, node, node, tfsoil_c=-3.0, extra_layers=ex)
lib = C.CDLL(os.path.join(HERE, "libemu_stage.so"))
s = b.as_struct()
key = np.zeros(n, np.uint64)
lib.stage_keys.argtypes = [C.POINTER(N.Batch), C.c_void_p]
lib.stage_keys(C.byref(s), key.ctypes.data)
order = np.argsort(key, kind="stable")
monthly = np.zeros((7, ny, 12, n), np.float32); yearly = np.zeros((5, ny, n), np.float32)
rawm = np.zeros((14, ny, 12, n), np.float32); flags = np.zeros(n, np.uint32); topo = np.zeros((16, n), np.int32)
lib.emu_run.argtypes = [C.POINTER(N.Batch)] + [C.c_void_p] * 5
lib.emu_run(C.byref(s), monthly.ctypes.data, yearly.ctypes.data, rawm.ctypes.data, flags.ctypes.data, topo.ctypes.data)
lib.stage_len.restype = C.c_longlong
def stream(c):
    L = lib.stage_len(int(c)); a = np.zeros(L, np.uint8); lib.stage_get(int(c), a.ctypes.data); return a.reshape(-1, 4)
SB = float(os.environ.get("SB_COST", "1.5"))     # a shared-base pair in units of one power
tot = dict(cur=0.0, seg=0.0, lanepow=0.0, sub=0.0, warps=0, skipped=0, cur_slots=0.0)
for w in range(n // 32):
    cells = order[w * 32:(w + 1) * 32]
    S = [stream(c) for c in cells]
    if len({len(x) for x in S}) != 1: tot["skipped"] += 1; continue
    A = np.stack(S).astype(np.float64)          # [32, L, 4] = tail single pows (of i-1), tail pairs (of i-1), cap of i, J
    J = A[0, :, 3]
    if not (A[:, :, 3] == J).all(): tot["skipped"] += 1; continue
    capraw = A[:, :, 2]
    dry = capraw == 4; cap = np.where(capraw >= 3, 3.0, 0.0)
    need = (cap > 0) & ~dry
    tot["cap_lane"] = tot.get("cap_lane", 0) + (cap > 0).sum(); tot["cap_lane_dry"] = tot.get("cap_lane_dry", 0) + dry.sum()
    anyc = (cap > 0).any(0); tot["cap_warp"] = tot.get("cap_warp", 0) + anyc.sum(); tot["cap_warp_skip"] = tot.get("cap_warp_skip", 0) + (anyc & ~need.any(0)).sum()
    hl = np.roll(A[:, :, 0], -1, axis=1); nw = np.roll(A[:, :, 1], -1, axis=1)
    hl[:, -1] = 0; nw[:, -1] = 0
    lane = cap + hl + SB * nw
    cur = cap.max(0) + hl.max(0) + SB * nw.max(0)
    seg = lane.max(0)
    # lanes free-running through a whole sub-step (sync where J restarts)
    starts = np.flatnonzero(J == J[0]); 
    cs = np.add.reduceat(lane, starts, axis=1)
    tot["cur"] += cur.sum(); tot["seg"] += seg.sum(); tot["sub"] += cs.max(0).sum(); tot["lanepow"] += lane.sum(); tot["warps"] += 1
print("warps %d (skipped %d)" % (tot["warps"], tot["skipped"]))
print("pow passes per warp, shared-base pair = %.1f pow:" % SB)
print("  today (stage by stage)            %.4g   lane efficiency %.1f %%" % (tot["cur"], 100 * tot["lanepow"] / (32 * tot["cur"])))
print("  state machine, sync per segment   %.4g   (%.1f %% of today)  lane efficiency %.1f %%" % (tot["seg"], 100 * tot["seg"] / tot["cur"], 100 * tot["lanepow"] / (32 * tot["seg"])))
print("  state machine, sync per sub-step  %.4g   (%.1f %% of today)  lane efficiency %.1f %%" % (tot["sub"], 100 * tot["sub"] / tot["cur"], 100 * tot["lanepow"] / (32 * tot["sub"])))
print("capillary limits: lane evaluations %d, of which storage at/below WP (skippable) %.1f %%; warp evaluations %d, all needing lanes skippable %.1f %%" % (
    tot["cap_lane"], 100.0 * tot["cap_lane_dry"] / tot["cap_lane"], tot["cap_warp"], 100.0 * tot["cap_warp_skip"] / tot["cap_warp"]))
