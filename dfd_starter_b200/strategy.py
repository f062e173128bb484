"""Strategy history and novelty with the reference's interface, evaluated in batches on the device
(SURVEY.md §8f row N3).

Mirrors `StrategyHandler` (strategy/strategy_handler.py:6-30), `SparseHistoryManager`
(strategy/sparse_history_manager.py:6-149) and `StrategyPoint` (strategy/strategy_point.py:6-39): same methods,
same replacement rule, same `strategy_tensor`.  What differs is how the arithmetic runs:

* `evaluate_strategies(zeta)`: all history policies are evaluated on the zeta frames in ONE batched forward (their
  parameter vectors staged as device rows) instead of one `set_trainable_flat` + forward per point;
* the pairwise distance table and every novelty come from `dfd_strategy_distances` (one launch each);
* `compute_novelty_members(idx, sign, sigma)` (new surface, used by the batched `Worker`) gives the novelty of every
  perturbed member at once: the reference computes `strategy_handler.compute_novelty(policy)` inside the per-member
  loop while the policy holds the perturbed parameters (worker/worker.py:53).

The novelty is shipped with every return but is commented out of the live objective
(learner/finite_differences.py:41-48); nothing here changes that.
"""
import numpy as np
import torch

from . import _lib
from .device import ptr

# utils/math_helpers.py:166-222, by function name (the drivers pass the function object itself)
DISTANCES = {
    "l2_dist": 0, "categorical_tvd": 1, "gaussian_wasserstein_dist_from_strategies": 2,
    "categorical_bhattacharrya_dist": 3, "gaussian_bhattacharrya_dist": 4,
}
_MAX_OBS_BYTES = 256 << 20          # zeta frames are replicated per member for the batched forward: bound the chunk


def distance_kind(fn):
    """A distance given by name, by kind number, or as a function whose __name__ is one of math_helpers' (so
    `math_helpers.categorical_tvd` from the reference's init_helper.py:12-30 can be passed as is)."""
    if isinstance(fn, int):
        return fn
    name = fn if isinstance(fn, str) else getattr(fn, "__name__", None)
    if name is None and fn is None:
        return DISTANCES["l2_dist"]                               # math_helpers.py:148-149 default
    if name not in DISTANCES:
        raise _lib.DfdError("unknown strategy distance %r (known: %s)" % (name, ", ".join(sorted(DISTANCES))))
    return DISTANCES[name]


def strategy_distances(ctx, a, b, kind, want_dists=True, want_min=True, exclude_diagonal=False):
    """a [n_a, Z, W], b [n_b, Z, W] device fp32 -> (dists [n_a, n_b] fp64 or None, row_min [n_a] fp64 or None)."""
    a, b = a.contiguous(), b.contiguous()
    n_a, Z, W = a.shape
    n_b = b.shape[0]
    dists = torch.empty(n_a, n_b, dtype=torch.float64, device=ctx.device) if want_dists else None
    rmin = torch.empty(n_a, dtype=torch.float64, device=ctx.device) if want_min else None
    _lib.check(ctx.lib.dfd_strategy_distances(ctx.handle, ptr(a), n_a, ptr(b), n_b, Z, W, int(kind), ptr(dists), ptr(rmin),
                                              1 if exclude_diagonal else 0, ctx.stream), "dfd_strategy_distances")
    return dists, rmin


class StrategyPoint(object):
    """strategy/strategy_point.py:6-39 (the parameter copy and the nearest / second-nearest bookkeeping)."""

    def __init__(self, policy, flat):
        self.flat = np.array(flat, dtype=np.float32)
        self.policy = policy
        self.strategy = None
        self.reset_dists()

    def add_dist(self, key, dist):
        if dist < self.closest[1]:
            self.second_closest = self.closest[:]
            self.closest = [key, dist]
        elif dist < self.second_closest[1] and key != self.closest[0]:
            self.second_closest = [key, dist]

    def reset_dists(self):
        self.closest = [None, np.inf]
        self.second_closest = [None, np.inf]


class SparseHistoryManager(object):
    def __init__(self, policy, strategy_distance_fn, max_history_size):
        self.policy = policy
        self.max_history_size = max_history_size
        self.worst_point_idx = 0
        self.strategy_points = []
        self.strategy_tensor = []
        self.zeta = []
        self.kind = distance_kind(strategy_distance_fn)
        self.known_dists = None                 # [n, n] fp64 host table (the reference keeps a dict of pairs)
        self._tensor_dev = None                 # [n, Z, W] on the device
        self._zeta_dev = None

    # ---- batched evaluation --------------------------------------------------------------------
    def _zeta_device(self, zeta):
        z = torch.as_tensor(np.asarray(zeta), dtype=torch.float32)
        return z.reshape((-1,) + self.policy._obs_shape()).to(self.policy.ctx.device)

    def _strategies_of_rows(self, flats):
        """[n, P] host parameter vectors -> their strategies on zeta, [n, Z, W] device."""
        from .noise_sources import RowTable
        pol = self.policy
        ctx = pol.ctx
        flats = np.ascontiguousarray(flats, dtype=np.float32)
        n = flats.shape[0]
        rt = RowTable(ctx, flats)
        zero = torch.zeros(pol.num_params, dtype=torch.float32, device=ctx.device)
        z = self._zeta_dev
        obs = z.unsqueeze(0).expand((n,) + tuple(z.shape)).contiguous()
        idx = torch.from_numpy(rt.idx).to(ctx.device)
        sign = torch.ones(n, dtype=torch.int8, device=ctx.device)
        out = pol.forward_members(idx, sign, obs, 1.0, theta=zero, table=rt)
        torch.cuda.current_stream(ctx.device).synchronize()      # rt is released on return
        return out

    def strategies_of_members(self, idx, sign, sigma):
        """Strategies of perturbed members theta + sign*sigma*table[idx] on zeta: [M, Z, W] device."""
        pol = self.policy
        dev = pol.ctx.device
        z = self._zeta_dev
        idx_d = torch.as_tensor(np.ascontiguousarray(idx), dtype=torch.int64).to(dev)
        sign_d = torch.as_tensor(np.ascontiguousarray(sign), dtype=torch.int8).to(dev)
        M = idx_d.shape[0]
        out = torch.empty(M, z.shape[0], pol.out_width, dtype=torch.float32, device=dev)
        step = max(1, int(_MAX_OBS_BYTES // max(1, z.numel() * 4)))
        for s in range(0, M, step):
            e = min(M, s + step)
            obs = z.unsqueeze(0).expand((e - s,) + tuple(z.shape)).contiguous()
            pol.forward_members(idx_d[s:e], sign_d[s:e], obs, sigma, out=out[s:e])
        return out

    # ---- reference interface -----------------------------------------------------------------------
    def submit_policy(self, policy):
        """sparse_history_manager.py:17-30."""
        point = StrategyPoint(self.policy, policy.get_trainable_flat())
        if len(self.strategy_points) >= self.max_history_size and self.zeta is not None and len(self.zeta) > 0:
            return self._replace_point(point)
        self.strategy_points.append(point)
        return None

    def evaluate_strategies(self, zeta):
        """:32-49: every history policy on zeta (one batched forward), then the distance table (one launch)."""
        self.zeta = zeta
        self._zeta_dev = self._zeta_device(zeta)
        pts = self.strategy_points
        if len(pts) == 0:
            self._tensor_dev, self.known_dists = None, np.zeros((0, 0))
            self.strategy_tensor = np.asarray([])
            return self.strategy_tensor
        self._tensor_dev = self._strategies_of_rows(np.stack([p.flat for p in pts]))
        self.strategy_tensor = self._tensor_dev.cpu().numpy()
        for p, s in zip(pts, self.strategy_tensor):
            p.strategy = s
        dists, _ = strategy_distances(self.policy.ctx, self._tensor_dev, self._tensor_dev, self.kind, want_min=False)
        self.known_dists = dists.cpu().numpy()
        self._update_strategy_point_dists()
        return self.strategy_tensor

    def _replace_point(self, point):
        """:72-109: the candidate replaces the least novel history point iff its novelty (distance to its nearest
        history strategy, the point about to be replaced included) exceeds that point's nearest-neighbour distance."""
        strat = self._strategies_of_rows(point.flat[None, :])
        dists, rmin = strategy_distances(self.policy.ctx, strat, self._tensor_dev, self.kind)
        dists = dists[0].cpu().numpy()
        novelty = float(rmin[0].item())
        idx = self.worst_point_idx
        current_worst = self.strategy_points[idx].closest[1]
        if novelty > current_worst or current_worst == np.inf:
            point.strategy = strat[0].cpu().numpy()
            self.strategy_points[idx] = point
            if idx < len(self.strategy_tensor):
                self.strategy_tensor[idx] = point.strategy
                self._tensor_dev[idx].copy_(strat[0])
            m = self.known_dists.shape[0]                        # pairs known since the last evaluate_strategies (:98-101)
            if idx < m:
                self.known_dists[idx, :] = dists[:m]
                self.known_dists[:, idx] = dists[:m]
            self._update_strategy_point_dists()
            return idx
        return -1

    def _update_strategy_point_dists(self):
        """:111-149 on the table: nearest and second-nearest neighbour of every point, then the point to replace next:
        of the closest pair, the one whose second-nearest neighbour is nearer.  Pairs are visited in the reference's
        dict order (other index ascending), so ties resolve the same way."""
        pts = self.strategy_points
        n = self.known_dists.shape[0]              # points added since the last evaluation have no pairs yet
        for p in pts:
            p.reset_dists()
        for i in range(n):
            row = self.known_dists[i]
            for j in range(n):
                if j != i:
                    pts[i].add_dist((min(i, j), max(i, j)), float(row[j]))
        worst_dist = np.inf
        for i in range(len(pts)):
            closest = pts[i].closest
            if closest[1] < worst_dist:
                worst_idx1 = i
                if closest[0] is None:
                    self.worst_point_idx = i
                    continue
                worst_idx2 = closest[0][1 - closest[0].index(i)]
                worst_dist = closest[1]
                if pts[worst_idx1].second_closest[1] < pts[worst_idx2].second_closest[1]:
                    self.worst_point_idx = worst_idx1
                else:
                    self.worst_point_idx = worst_idx2


class StrategyHandler(object):
    def __init__(self, policy, strategy_distance_fn, max_history_size=200):
        self.strategy_history_manager = SparseHistoryManager(policy, strategy_distance_fn, max_history_size)
        self.strategy_tensor = np.zeros(0)
        self.policy = policy
        self.zeta = None
        self.max_history_size = max_history_size
        self.strategy_distance_fn = strategy_distance_fn

    def add_policy(self, policy):
        self.strategy_history_manager.submit_policy(policy)

    def set_zeta(self, zeta):
        if zeta is None or len(zeta) == 0:
            return
        self.zeta = zeta
        self.strategy_tensor = self.strategy_history_manager.evaluate_strategies(zeta)

    def _ready(self):
        return not (self.zeta is None or len(self.zeta) == 0 or self.strategy_tensor is None or len(self.strategy_tensor) < 2)

    def compute_novelty(self, policy):
        """strategy_handler.py:25-30: novelty of `policy`'s current parameters."""
        if not self._ready():
            return 0
        m = self.strategy_history_manager
        strat = m._strategies_of_rows(np.asarray(policy.get_trainable_flat(), dtype=np.float32)[None, :])
        _, rmin = strategy_distances(self.policy.ctx, strat, m._tensor_dev, m.kind, want_dists=False)
        return float(rmin[0].item())

    def compute_novelty_members(self, idx, sign, sigma):
        """Novelty of every member theta + sign*sigma*table[idx] (sign 0: unperturbed): float64 [M]."""
        if not self._ready():
            return np.zeros(len(idx))
        m = self.strategy_history_manager
        strat = m.strategies_of_members(idx, sign, sigma)
        _, rmin = strategy_distances(self.policy.ctx, strat, m._tensor_dev, m.kind, want_dists=False)
        return rmin.cpu().numpy()

