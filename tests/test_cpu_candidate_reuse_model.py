"""The invariant behind the broad phase's candidate reuse (csrc/k_broad.cu: bpCheckKernel / bpDecideKernel / bpFilterKernel), as a small numpy
model on the CPU: the pair set is a function of collider indices and AABBs only, so a candidate list built from boxes grown by a margin,
filtered IN ORDER by the exact inclusive predicate on the true boxes, equals the canonical pair list for as long as every true box lies
inside its grown box -- and the containment check is what notices when it no longer does (moved too far, grew by tumbling, teleported,
a padding slot that became a collider). The CUDA implementation is compared bit for bit with a context that rebuilds in every step in the
-m gpu suite (tests/test_gpu_parity_scale.py, tests/mgpu_reuse_check.py); this pins the argument itself."""
import numpy as np


def exact_pairs(mn, mx):
    """canonical list: a < b, sorted by (a, b), inclusive overlap (bounding_volumes.h:352-358); inverted boxes (padding slots) are in no pair"""
    n = len(mn); valid = mn[:, 0] <= mx[:, 0]
    out = []
    for a in range(n):
        if not valid[a]: continue
        hit = np.all(mx[a] >= mn[a + 1:], axis=1) & np.all(mn[a] <= mx[a + 1:], axis=1) & valid[a + 1:]
        out.extend((a, a + 1 + int(j)) for j in np.nonzero(hit)[0])
    return out


class ReuseModel:
    def __init__(self, margin):
        self.margin = margin; self.gmn = self.gmx = None; self.cand = None; self.rebuilds = 0

    def step(self, mn, mx):
        rebuild = self.cand is None or len(mn) != len(self.gmn)
        if not rebuild:
            now_pad = mn[:, 0] > mx[:, 0]; was_pad = self.gmn[:, 0] > self.gmx[:, 0]
            inside = np.all(mn >= self.gmn, axis=1) & np.all(mx <= self.gmx, axis=1)
            rebuild = bool(np.any(~now_pad & (was_pad | ~inside)))
        if rebuild:
            pad = mn[:, 0] > mx[:, 0]
            self.gmn = np.where(pad[:, None], mn, mn - self.margin); self.gmx = np.where(pad[:, None], mx, mx + self.margin)
            self.cand = exact_pairs(self.gmn, self.gmx); self.rebuilds += 1
        keep = [(a, b) for a, b in self.cand
                if mn[a, 0] <= mx[a, 0] and mn[b, 0] <= mx[b, 0] and np.all(mx[a] >= mn[b]) and np.all(mn[a] <= mx[b])]
        return keep


def test_filtered_candidates_equal_the_exact_pair_list_while_boxes_stay_inside_their_grown_boxes():
    rng = np.random.default_rng(11)
    n = 160
    centre = rng.uniform(0, 6, size=(n, 3)); half = rng.uniform(0.2, 0.6, size=(n, 3))
    vel = rng.normal(0, 0.002, size=(n, 3)); vel[::17] *= 5.0                           # a few fast movers among slow ones
    model = ReuseModel(margin=0.05)
    pad = np.zeros(n, dtype=bool); pad[-6:] = True                                     # halo padding slots: inverted boxes
    for step in range(60):
        if step == 30: centre[5] += np.array([2.0, -1.0, 0.5])                         # a teleported body
        if step == 40: pad[-3:] = False                                                # padding slots become colliders (a ghost arrives)
        if step == 50: half[9] *= 1.4                                                  # a box that grows (a tumbling OBB)
        centre += vel
        mn = centre - half; mx = centre + half
        mn[pad] = 3.0e38; mx[pad] = -3.0e38
        assert model.step(mn, mx) == exact_pairs(mn, mx), "step %d" % step
    assert 3 <= model.rebuilds < 60                                                    # reused in between, rebuilt when it had to be


def test_a_zero_margin_list_is_only_valid_for_its_own_step():
    rng = np.random.default_rng(3)
    centre = rng.uniform(0, 3, size=(60, 3)); half = np.full((60, 3), 0.3)
    model = ReuseModel(margin=0.0)
    for step in range(5):
        centre += rng.normal(0, 0.01, size=centre.shape)
        mn, mx = centre - half, centre + half
        assert model.step(mn, mx) == exact_pairs(mn, mx)
    assert model.rebuilds == 5          # every box has moved out of its (ungrown) box: the device never even tries (bpDecideKernel: margin 0 -> rebuild)
