"""Verbose GPU-vs-oracle check used during development (run under gpurun)."""
import sys, time, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from superodom_b200 import synth, api
from oracle import oracle as O

name = sys.argv[1] if len(sys.argv) > 1 else "cfg1"
c = synth.make_case(name)
cfg = c["cfg"]
print("case", name, "map", c["map_xyzi"].shape, "scan", c["scan_xyzi"].shape, flush=True)
ctx = api.Context(max_map_points=max(1 << 20, len(c["map_xyzi"]) + 1024), max_scan_points=262144, plane_res=cfg["plane_res"])
t = time.time(); ctx.map_set_points(c["map_xyzi"]); print("gpu map build %.1f ms" % ((time.time() - t) * 1e3))
om = O.OracleMap(c["map_xyzi"])

# k-NN
rng = np.random.default_rng(0)
q = c["map_xyzi"][::17, :3] + rng.normal(0, 0.1, size=c["map_xyzi"][::17, :3].shape).astype(np.float32)
gi, gd = ctx.knn(q, 5, 0.0)
oi, od, of = om.knn(q, 5, 0)
gi64 = gi.astype(np.int64); gi64[gi == 0xFFFFFFFF] = -1
print("knn exact: idx equal", (gi64 == oi).all(), "mismatch rows", int((gi64 != oi).any(1).sum()), "d2 equal", (gd == od).all())
gi, gd = ctx.knn(q, 5, 3 * cfg["plane_res"])
print("knn bounded: found5", int((gi[:, 4] != 0xFFFFFFFF).sum()), "of", len(q))

# correspondences at the prior pose
cap = cfg["max_surface_features"]
gc, gho, ghr = ctx.correspond(c["scan_xyzi"], c["pose_prior"], cap)
oc, oho, ohr = om.correspond(c["scan_xyzi"], c["pose_prior"], cfg["plane_res"], cap, 0)
ost = oc["status"].copy(); ost[ost < 0] = 255
print("status equal:", (gc["status"] == ost).all(), "diff", int((gc["status"] != ost).sum()))
print("hist obs", gho, oho, "rej", ghr, ohr)
ok = (ost == 0) & (gc["status"] == 0)
print("n_ok", ok.sum())
print("nn equal (ok rows):", (gc["nn"][ok].astype(np.int64) == oc["nn"][ok]).all())
for f in ("n", "d", "w"):
    a, b = gc[f][ok], oc[f][ok]
    print(f, "max rel diff", np.max(np.abs(a - b) / np.maximum(np.abs(b), 1e-300)))
print("obs equal:", (gc["obs"][ok] == oc["obs"][ok]).all(), int((gc["obs"][ok] != oc["obs"][ok]).any(1).sum()))
H, g, cost = ctx.evaluate(c["pose_prior"])
oH, og, ocost, nok = O.evaluate(oc, c["pose_prior"], cfg["plane_res"])
print("H rel", np.abs(H - oH).max() / np.abs(oH).max(), "g rel", np.abs(g - og).max() / np.abs(og).max(), "cost rel", abs(cost - ocost) / ocost)

# full registration
for trial in range(2):
    t = time.time()
    r = ctx.register(c["scan_xyzi"], c["pose_prior"], cfg["max_iterations"], cap)
    dt = (time.time() - t) * 1e3
    print("gpu register: status", r.status, "iters", r.n_iterations, "time_ms(dev) %.3f total %.3f wall %.3f" % (r.time_ms, r.time_total_ms, dt))
ro = om.register(c["scan_xyzi"], c["pose_prior"], cfg["plane_res"], cfg["max_iterations"], cap, knn_mode=0)
pg, po = np.array(r.pose), np.array(ro.pose)
print("pose gpu   ", pg)
print("pose oracle", po)
print("dpos %.3e  dquat %.3e" % (np.abs(pg[:3] - po[:3]).max(), min(np.abs(pg[3:] - po[3:]).max(), np.abs(pg[3:] + po[3:]).max())))
print("iters", r.n_iterations, ro.n_iterations, "nsurf", list(r.iter_n_surf[:r.n_iterations]), list(ro.iter_n_surf[:ro.n_iterations]))
print("lm steps", list(r.iter_lm_steps[:r.n_iterations]), list(ro.iter_lm_steps[:ro.n_iterations]), "succ", list(r.iter_lm_successful[:r.n_iterations]), list(ro.iter_lm_successful[:ro.n_iterations]),
      "term", list(r.iter_lm_termination[:r.n_iterations]), list(ro.iter_lm_termination[:ro.n_iterations]))
print("cost", np.array(r.iter_cost[:r.n_iterations]), np.array(ro.iter_cost[:ro.n_iterations]))
cg, co = np.array(r.cov).reshape(6, 6), np.array(ro.cov).reshape(6, 6)
print("cov rel", np.abs(cg - co).max() / np.abs(co).max(), "pos_err", r.pos_err, ro.pos_err, "inv_cond", r.pos_inv_cond, ro.pos_inv_cond, "ori", r.ori_err_deg, ro.ori_err_deg)
print("oracle time_ms %.1f (knn %.1f)" % (ro.time_ms, ro.time_knn_ms))
print("launches", ctx.kernel_launches())
