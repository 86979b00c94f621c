"""CPU oracle for LOST object discovery — TEST INFRASTRUCTURE ONLY.

numpy restatement of `object_discovery.py:23-134` (lost / patch_scoring / detect_box) of the
reference.  Only tests/, smoke() and bench.py's CPU-baseline legs may import it.

Parity status: PINNED against the unmodified reference run in the build container
(`tests/golden/make_golden.py` -> `tests/golden/lost_*.npz`), with the reference's unstable
`torch.argsort` pinned to `stable=True` ("lowest patch index first among equal degrees",
SURVEY §8c) — that is the tie policy this oracle and the CUDA path implement.

The Gram matrix is computed in fp32 like the reference (`feats @ feats.T`); the fp64 variant is
used by tests to decide which near-zero entries may legitimately differ in sign.
"""
import numpy as np

F32 = np.float32


def gram(feats, dtype=F32):
    """object_discovery.py:39: A = (feats @ feats.transpose(1, 2)).squeeze(); feats is [N, d]."""
    f = np.asarray(feats, dtype)
    return f @ f.T


def patch_scoring(A, threshold=0.0):
    """object_discovery.py:72-90.  Returns (sel, cent): patches by ascending degree (stable), and
    cent = -degree as fp32, degree_i = #{j != i : A_ij > threshold}."""
    A = np.array(A, dtype=F32, copy=True)
    np.fill_diagonal(A, 0)                                   # :80
    A[A < 0] = 0                                             # :83
    cent = -(A > threshold).sum(axis=1).astype(F32)          # :87
    sel = np.argsort(-cent, kind="stable")                   # :88 argsort(cent, descending=True), ties pinned
    return sel, cent


def label_component(binary, seed_rc):
    """4-connected component of `binary` containing seed_rc (scipy.ndimage.label's default
    cross structure, object_discovery.py:104-107).  Returns a bool map, or None when the seed is
    background."""
    h, w = binary.shape
    r0, c0 = seed_rc
    if not binary[r0, c0]:
        return None
    comp = np.zeros_like(binary, dtype=bool)
    stack = [(r0, c0)]
    comp[r0, c0] = True
    while stack:
        r, c = stack.pop()
        for dr, dc in ((1, 0), (-1, 0), (0, 1), (0, -1)):
            rr, cc = r + dr, c + dc
            if 0 <= rr < h and 0 <= cc < w and binary[rr, cc] and not comp[rr, cc]:
                comp[rr, cc] = True
                stack.append((rr, cc))
    return comp


def detect_box(M, seed, dims, initial_im_size=None, scales=None):
    """object_discovery.py:93-134.  M: [N] correlation of every patch with the expanded seed."""
    w_featmap, h_featmap = dims
    correl = np.asarray(M, F32).reshape(w_featmap, h_featmap)            # :101
    seed_rc = np.unravel_index(int(seed), (w_featmap, h_featmap))        # :107
    comp = label_component(correl > 0.0, seed_rc)                         # :104
    if comp is None:
        raise ValueError("The seed is in the background component.")     # :110-111
    rows, cols = np.where(comp)                                           # :114
    ymin, ymax = rows.min(), rows.max() + 1                               # :116
    xmin, xmax = cols.min(), cols.max() + 1                               # :117
    pred = [scales[1] * xmin, scales[0] * ymin, scales[1] * xmax, scales[0] * ymax]   # :120-123
    if initial_im_size:                                                   # :126-128
        pred[2] = min(pred[2], initial_im_size[1])
        pred[3] = min(pred[3], initial_im_size[0])
    return pred, [ymin, xmin, ymax, xmax]


def lost(feats, dims, scales, init_image_size, k_patches=100, A=None):
    """object_discovery.py:23-69 for one image.  feats: [N, d] (the reference's [1, N, d] squeezed).
    Returns (pred[4], A[N,N], scores[N], seed)."""
    if A is None:
        A = gram(feats)
    sel, scores = patch_scoring(A)                                        # :54
    seed = int(sel[0])                                                    # :57
    potentials = sel[:k_patches]                                          # :60
    similars = potentials[A[seed, potentials] > 0.0]                      # :61
    M = A[similars, :].sum(axis=0, dtype=F32)                             # :62
    pred, _ = detect_box(M, seed, dims, scales=scales, initial_im_size=init_image_size[1:])   # :65-67
    return np.asarray(pred), A, scores, seed


def planted_object_feats(rng, grid=(30, 30), d=384, rows=(8, 18), cols=(10, 22), noise=1.0):
    """Synthetic image with one object: background direction + object direction +
    noise, mean-centred.  Few degree ties, recovers the planted box."""
    h, w = grid
    bg = rng.standard_normal(d).astype(F32)
    obj = rng.standard_normal(d).astype(F32)
    base = np.tile(bg, (h * w, 1))
    grid_mask = np.zeros((h, w), dtype=bool)
    grid_mask[rows[0]:rows[1], cols[0]:cols[1]] = True
    base[grid_mask.reshape(-1)] = obj
    feats = base + noise * rng.standard_normal((h * w, d)).astype(F32)
    feats = feats - feats.mean(axis=0, keepdims=True)
    return feats.astype(F32)


def lost_from_degrees(feats, degree, dims, scales, init_image_size, k_patches=100, ulps=64.0):
    """Checker for implementations whose Gram entries differ from the reference's in rounding only.

    Takes the DEGREES the implementation produced (they may differ from the reference's where a Gram entry is within
    rounding of zero; the caller checks that separately against the fp64 Gram) and replays the rest of
    object_discovery.py:57-67 in float64: seed = stable argmin, potentials, similars (A[seed, p] > 0), M = sum of the
    similar rows, flood fill, box.  Returns (seed, pred or None when the seed is background, decidable): `decidable` is
    False when some A[seed, p] or some M_j lies within `ulps` fp32 ulps of zero relative to its natural scale
    (|k_seed||k_p|, resp. |k_j| sqrt(sum_s |k_s|^2): the rounding of a length-d fp32 dot product / of the reference's own
    fp32 row sum), i.e. when fp32 implementations may legitimately disagree on the sign.  seed is always binding."""
    eps = float(np.finfo(F32).eps)
    f = np.asarray(feats, np.float64)
    A = f @ f.T
    norms = np.sqrt(np.maximum(np.diag(A), 0.0))
    sel = np.argsort(np.asarray(degree, np.int64), kind="stable")
    seed = int(sel[0])
    pot = sel[:k_patches]
    a = A[seed, pot]
    decidable = bool(np.all(np.abs(a) > ulps * eps * norms[seed] * norms[pot]))
    sim = pot[a > 0.0]
    M = A[sim, :].sum(axis=0)
    tol = ulps * eps * norms * np.sqrt(np.sum(norms[sim] ** 2))
    decidable = decidable and bool(np.all(np.abs(M) > tol))
    try:
        pred, _ = detect_box(M, seed, dims, scales=scales, initial_im_size=init_image_size[1:])
    except ValueError:
        pred = None
    return seed, pred, decidable
