"""CPU: the inequalities the tensor-core filter (csrc/nda_tc.cuh) rests on, checked numerically in float64
and with the fp16 rounding of the operands emulated in numpy.

  (1) chord bound      S = sum_k (c_k/pi)^2 sin^2(pi w_k / c_k) <= r^2 = sum_k w_k^2   (w = minimum image)
  (2) dot identity     phi(a).phi(b) = P - S/2,  P = sum_k (c_k / 2 pi)^2
  (3) rounding slack   |dot(fp16 operands, A compensated) - dot_exact| <= E = P (2^-11 * 1.01 + 2^-18)
  (4) candidate rule   r^2 <= thr  =>  dot_fp16 >= P - thr/2 - E            (no in-range pair is missed)
  (5) SURE rule        dot_exact >= P - rs^2 / (2 kappa)  =>  r^2 <= rs^2,  kappa = max_k (asin u_k / u_k)^2
"""
import numpy as np

TWO_PI = 2.0 * np.pi


def embed(x, c):
    """(n, 3) coordinates -> (n, 6) exact embedding in float64"""
    th = TWO_PI * x / c
    rho = c / TWO_PI
    return np.stack([rho * np.cos(th), rho * np.sin(th)], -1).reshape(x.shape[0], 6)


def min_image(a, b, c):
    d = a[:, None, :] - b[None, :, :]
    return d - c * np.round(d / c)


def setup(seed, n=400, far=3.0):
    rng = np.random.default_rng(seed)
    c = rng.uniform(25.0, 90.0, 3)
    a = rng.uniform(-far, far + 1.0, (n, 3)) * c          # several cells away from the origin
    b = rng.uniform(-far, far + 1.0, (n + 17, 3)) * c
    return c, a, b


def test_chord_bound_and_dot_identity():
    for seed in range(5):
        c, a, b = setup(seed)
        w = min_image(a, b, c)
        r2 = (w ** 2).sum(-1)
        S = (((c / np.pi) * np.sin(np.pi * w / c)) ** 2).sum(-1)
        assert np.all(S <= r2 * (1 + 1e-12) + 1e-12)
        P = ((c / TWO_PI) ** 2).sum()
        dot = embed(a, c) @ embed(b, c).T
        assert np.max(np.abs(dot - (P - S / 2))) < 1e-9 * P


def fp16_operands(c, a, b):
    ea, eb = embed(a, c), embed(b, c)
    a_hi = ea.astype(np.float16).astype(np.float64)
    a_lo = (ea - a_hi).astype(np.float16).astype(np.float64)
    b_hi = eb.astype(np.float16).astype(np.float64)
    return ea, eb, a_hi, a_lo, b_hi


def test_rounding_slack_and_candidate_rule():
    for seed in range(5):
        c, a, b = setup(100 + seed)
        ea, eb, a_hi, a_lo, b_hi = fp16_operands(c, a, b)
        P = ((c / TWO_PI) ** 2).sum()
        E = P * (2.0 ** -11 * 1.01 + 2.0 ** -18)
        dot_exact = ea @ eb.T
        dot16 = (a_hi + a_lo) @ b_hi.T                     # what the MMA sums exactly (K = 16: [hi | lo] x [hi | hi])
        assert np.max(np.abs(dot16 - dot_exact)) <= E
        w = min_image(a, b, c)
        r2 = (w ** 2).sum(-1)
        for R in (1.0, 3.0, 7.5, 0.27 * c.min()):
            thr = R * R
            in_range = r2 <= thr
            assert np.all(dot16[in_range] >= P - thr / 2 - E)
            # and the filter is selective: for ranges well below the cell size, what it lets through lies
            # within a thin shell around the range (chord deficit (x / sin x)^2 with x = pi r / c)
            cand = dot16 >= P - thr / 2 - 1.25 * E
            if R <= 0.1 * c.min() and cand.any():
                assert np.sqrt(r2[cand]).max() <= np.sqrt(thr + 4.5 * E) * 1.03


def test_sure_rule():
    rng = np.random.default_rng(7)
    for seed in range(5):
        c, a, _ = setup(200 + seed, n=300)
        for rs in (1.0, 3.0, 0.2 * c.min(), 0.25 * c.min()):   # the kernel enables the rule for pi rs / c_min < 0.9
            # partners on a shell around rs, some inside, some outside, through periodic images
            dirs = rng.normal(size=(a.shape[0], 3))
            dirs /= np.linalg.norm(dirs, axis=-1, keepdims=True)
            rad = rs * rng.uniform(0.9, 1.1, (a.shape[0], 1))
            b = a + dirs * rad + rng.integers(-2, 3, a.shape) * c
            w = min_image(a, b, c)
            r2 = (w ** 2).sum(-1)
            dot = embed(a, c) @ embed(b, c).T
            P = ((c / TWO_PI) ** 2).sum()
            u = np.pi * rs / c
            kappa = ((np.arcsin(u) / u) ** 2).max()
            sure = dot >= P - rs * rs / (2 * kappa)
            assert np.all(r2[sure] <= rs * rs * (1 + 1e-9))
            assert sure.sum() > 0                           # the rule does fire (it decides ~9 of 10 found atoms)
            # the gap between the two thresholds is a thin shell: every pair closer than rs / sqrt(kappa) is sure
            assert np.all(sure[r2 <= rs * rs / kappa * (1 - 1e-9)])
