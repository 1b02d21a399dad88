"""A second, independent restatement of the Ceres solve the registration runs per ICP iteration (LidarSlam.cpp:213-240), in numpy,
cross-checked against oracle/so_oracle.cpp on random problems.

Ceres is not installable here, so the oracle's trust-region loop is a restatement of Ceres 2.0.0's published algorithm
(trust_region_minimizer.cc, levenberg_marquardt_strategy.cc, trust_region_step_evaluator.cc, corrector.cc, loss_function.cc) and so
is this file -- written separately, with numpy's LAPACK least squares instead of the oracle's hand-written Householder QR, its own
quaternion algebra and its own loop structure.  Two restatements that agree on the CONTROL FLOW (LM iterations, successful steps,
termination reason: what drives the outer ICP loop through `summary.num_successful_steps == 1`, LidarSlam.cpp:141) and on the
minimiser to 1e-9 on a hundred random problems shrink the room for a transcription slip in either; they do not replace a golden
vector from Ceres itself, which the reference does not ship ("parity unpinned" for this part).
"""
import numpy as np

FUNCTION_TOL, GRADIENT_TOL, PARAMETER_TOL = 1e-6, 1e-10, 1e-8
MIN_REL_DECREASE, MIN_LM_DIAG, MAX_LM_DIAG = 1e-3, 1e-6, 1e32
MAX_RADIUS, MIN_RADIUS = 1e16, 1e-32


def _qmul(a, b):              # xyzw, Hamilton
    ax, ay, az, aw = a
    bx, by, bz, bw = b
    return np.array([aw * bx + ax * bw + ay * bz - az * by, aw * by - ax * bz + ay * bw + az * bx,
                     aw * bz + ax * by - ay * bx + az * bw, aw * bw - ax * bx - ay * by - az * bz])


def _rot(q):
    x, y, z, w = q
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                     [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                     [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])


def _plus(x, d):              # PoseLocalParameterization::Plus (pose_local_parameterization.cpp:7-22)
    q = _qmul(x[3:], np.array([d[3] / 2, d[4] / 2, d[5] / 2, 1.0]))
    return np.concatenate([x[:3] + d[:3], q / np.linalg.norm(q)])


class Problem:
    """n point-to-plane blocks {p, n, d, w} under ScaledLoss(TukeyLoss(a), w) (LidarSlam.cpp:268-274)."""

    def __init__(self, P, N, D, W, a2):
        self.P, self.N, self.D, self.W, self.a2 = P, N, D, W, a2

    def _residuals(self, x):
        R = _rot(x[3:])
        return (self.N * (self.P @ R.T + x[:3])).sum(1) + self.D, R

    def _rho(self, s):
        inside = s <= self.a2
        v = np.where(inside, 1.0 - s / self.a2, 0.0)
        rho0 = np.where(inside, self.a2 / 6.0 * (1.0 - v ** 3), self.a2 / 6.0) * self.W
        rho1 = np.where(inside, 0.5 * v * v, 0.0) * self.W
        return rho0, rho1

    def cost(self, x):
        r, _ = self._residuals(x)
        return 0.5 * self._rho(r * r)[0].sum()

    def linearise(self, x):
        """cost, corrected residuals, corrected local Jacobian [n, 6] (Corrector with rho'' <= 0: both scaled by sqrt(rho'))."""
        r, R = self._residuals(x)
        rho0, rho1 = self._rho(r * r)
        a = self.N @ R                                     # rows: R^T n
        J = np.concatenate([self.N, np.cross(self.P, a)], 1)      # [n^T, -n^T R [p]x] = [n, p x (R^T n)]
        sc = np.sqrt(rho1)
        return 0.5 * rho0.sum(), r * sc, J * sc[:, None]


def solve(prob, x0, max_iterations=4):
    """-> (x, iterations, successful_steps, unsuccessful_steps, termination)  termination codes as oracle/so_oracle.cpp."""
    x = np.array(x0, float)
    cost, res, J = prob.linearise(x)
    scale = 1.0 / (1.0 + np.linalg.norm(J, axis=0))        # Jacobi scaling, fixed at iteration 0
    Js = J * scale
    g = J.T @ res

    def gmax_of(xx, gg):
        return np.abs(xx - _plus(xx, -gg)).max()
    gmax = gmax_of(x, g)
    radius, decrease, reuse = 1e4, 2.0, False
    diag = None
    it = succ = unsucc = invalid = 0
    step_ok = False
    while True:
        if step_ok:
            succ += 1
        else:
            unsucc += 1
        if it >= max_iterations:
            return x, it, succ, unsucc, 0
        if gmax <= GRADIENT_TOL:
            return x, it, succ, unsucc, 1
        if radius <= MIN_RADIUS:
            return x, it, succ, unsucc, 4
        it += 1
        step_ok = False
        if not reuse:
            diag = np.clip((Js * Js).sum(0), MIN_LM_DIAG, MAX_LM_DIAG)
        reuse = True
        A = np.concatenate([Js, np.diag(np.sqrt(diag / radius))], 0)
        b = np.concatenate([res, np.zeros(6)])
        y = np.linalg.lstsq(A, b, rcond=None)[0]
        step = -y
        mr = Js @ step
        model_change = -(mr * (res + mr / 2.0)).sum()
        if not (np.isfinite(step).all() and model_change > 0.0):
            invalid += 1
            if invalid >= 5:
                return x, it, succ, unsucc, 5
            radius *= 0.5
            continue
        invalid = 0
        cand = _plus(x, step * scale)
        cand_cost = prob.cost(cand)
        if np.linalg.norm(x - cand) <= PARAMETER_TOL * (np.linalg.norm(x) + PARAMETER_TOL):
            return x, it, succ, unsucc, 2
        change = cost - cand_cost
        if abs(change) <= FUNCTION_TOL * cost:
            return x, it, succ, unsucc, 3
        rel = change / model_change
        if rel > MIN_REL_DECREASE:
            x = cand
            cost, res, J = prob.linearise(x)
            Js = J * scale
            gmax = gmax_of(x, J.T @ res)
            step_ok = True
            radius = min(MAX_RADIUS, radius / max(1.0 / 3.0, 1.0 - (2.0 * rel - 1.0) ** 3))
            decrease, reuse = 2.0, False
        else:
            radius /= decrease
            decrease *= 2.0
            reuse = True


def _random_problem(rng, n):
    """Points on a few planes seen from a sensor, a ground-truth pose, and correspondences {p (sensor frame), n, d, w}."""
    truth = np.concatenate([rng.uniform(-5, 5, 3), (lambda q: q / np.linalg.norm(q))(np.concatenate([rng.normal(0, 0.15, 3), [1.0]]))])
    R = _rot(truth[3:])
    normals = np.linalg.qr(rng.normal(size=(3, 3)))[0]
    k = rng.integers(0, 3, n)
    Nrm = normals[k] * rng.choice([-1.0, 1.0], n)[:, None]
    world = rng.uniform(-20, 20, (n, 3))
    offs = rng.uniform(2, 15, 3)[k]
    world -= ((world * Nrm).sum(1) - offs)[:, None] * Nrm                 # onto the plane n.x = offs
    D = -offs + rng.normal(0, 0.01, n)                                     # a little plane noise
    P = (world - truth[:3]) @ R                                            # sensor frame: R^T (x - t)
    W = rng.uniform(0.5, 1.0, n)
    return truth, P, Nrm, D, W


def test_numpy_and_cpp_restatements_of_the_ceres_solve_agree():
    from oracle import oracle as O
    O.build()
    rng = np.random.default_rng(42)
    plane_res = 0.2
    a = float(np.sqrt(np.float32(3 * np.float32(plane_res))))             # TukeyLoss(std::sqrt(3 * planeRes_)): float sqrt (LidarSlam.cpp:271)
    a2 = a * a
    seen_terms, seen_unsucc = set(), 0
    for trial in range(100):
        n = int(rng.integers(40, 400))
        truth, P, Nrm, D, W = _random_problem(rng, n)
        # every fourth start is far off, every fifth very far (Tukey saturation, rejected steps, radius shrinking)
        dt, dr = (2.5, 0.6) if trial % 5 == 4 else ((0.6, 0.12) if trial % 4 == 3 else (0.08, 0.015))
        start = _plus(truth, np.concatenate([rng.uniform(-1, 1, 3) * dt, rng.uniform(-1, 1, 3) * dr]))
        corr = np.zeros(n, O.CORR_DTYPE)
        corr["p"], corr["n"], corr["d"], corr["w"], corr["status"] = P, Nrm, D, W, 0
        for lm in (1, 2, 4, 12):
            xo, so = O.solve(corr, start, plane_res, lm)
            xn, it, succ, unsucc, term = solve(Problem(P, Nrm, D, W, a2), start, lm)
            assert (it, succ, unsucc, term) == (so["iterations"], so["successful"], so["unsuccessful"], so["termination"]), (trial, lm, (it, succ, unsucc, term), so)
            assert np.abs(xn - xo).max() <= 1e-9, (trial, lm, np.abs(xn - xo).max())
            seen_terms.add(term)
            seen_unsucc += unsucc > 1
    # the sample exercises the max-iterations exit and the function / parameter tolerance exits (rejected steps do not occur with this
    # bounded loss from any start tried here: the rejection branch is only cross-read, not cross-run)
    assert {0, 2, 3} <= seen_terms, (seen_terms, seen_unsucc)
