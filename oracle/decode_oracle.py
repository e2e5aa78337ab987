"""numpy restatement of the heat-map -> key-point decode (TEST INFRASTRUCTURE).

Follows reference src/model_utils.py:10-51:
  * ``argmax_ind`` :10-16      flat ``np.argmax`` (first occurrence in row-major order; a NaN
                               compares as the maximum, so the first NaN wins) -> (row, col)
  * ``weighted_max_loc`` :18-36  5x5 window around the peak clipped to the map (:24-29);
                               centroid of (index + 0.5) weighted by the window's column / row
                               sums, divided by the window sum (:31-32); scaled by
                               ``target_size[0] / W`` for x and ``target_size[1] / H`` for y (:33-34)
  * ``get_keypoints_from_heatmaps(_batch)`` :38-51  loop over key-points / images.

The floating-point evaluation ORDER is part of the contract (the CUDA kernel reproduces it so
that the refined coordinates are bit-identical, not just the indices).  Probed against
numpy 2.3 on this image:
  * window column sums / row sums: float32, sequential over the reduced axis;
  * window total: float32, numpy's pairwise sum over the row-major flattened window
    (8 running partials over blocks of 8, combined as ((r0+r1)+(r2+r3))+((r4+r5)+(r6+r7)),
    then the <8 leftover values added sequentially) -- windows always hold 9..25 values;
  * weighted sums: float64 products of (index+0.5) with the float32 sums, added sequentially;
  * division by the float32 total promoted to float64, then ``/ size * target`` in float64.
No guard for a zero / negative window sum: inf / nan / out-of-window results are reproduced.
"""
from __future__ import annotations

import numpy as np

F32 = np.float32
PAD = 2


def first_argmax_loop(flat):
    """Definitional form: index of the first maximum, NaN treated as larger than everything."""
    best = 0
    bv = flat[0]
    if bv != bv:
        return 0
    for i in range(1, flat.size):
        v = flat[i]
        if v != v:
            return i
        if v > bv:
            bv, best = v, i
    return best


def first_argmax(flat):
    """Same rule as ``first_argmax_loop`` on numpy primitives (fast path for big batches)."""
    nan = np.flatnonzero(flat != flat)
    if nan.size:
        return int(nan[0])
    return int(np.flatnonzero(flat == flat.max())[0])


def _pairwise8_f32(vals):
    n = len(vals)
    if n < 8:
        acc = F32(-0.0)
        for v in vals:
            acc = F32(acc + v)
        return acc
    r = [F32(v) for v in vals[:8]]
    i = 8
    while i < n - (n % 8):
        for j in range(8):
            r[j] = F32(r[j] + vals[i + j])
        i += 8
    res = F32(F32(F32(r[0] + r[1]) + F32(r[2] + r[3])) + F32(F32(r[4] + r[5]) + F32(r[6] + r[7])))
    while i < n:
        res = F32(res + vals[i])
        i += 1
    return res


def decode_map(hm, target_size=(224, 224)):
    """One [H,W] float32 map -> (row, col, x, y) with x,y float64."""
    hm = np.asarray(hm, dtype=F32)
    H, W = hm.shape
    flat = first_argmax(hm.reshape(-1))
    cy, cx = divmod(flat, W)
    x0, x1 = max(0, cx - PAD), min(W, cx + PAD + 1)
    y0, y1 = max(0, cy - PAD), min(H, cy + PAD + 1)
    with np.errstate(all="ignore"):
        col = []
        for x in range(x0, x1):
            a = hm[y0, x]
            for y in range(y0 + 1, y1):
                a = F32(a + hm[y, x])
            col.append(a)
        row = []
        for y in range(y0, y1):
            a = hm[y, x0]
            for x in range(x0 + 1, x1):
                a = F32(a + hm[y, x])
            row.append(a)
        total = _pairwise8_f32([hm[y, x] for y in range(y0, y1) for x in range(x0, x1)])
        sx = np.float64(-0.0)
        for j, x in enumerate(range(x0, x1)):
            sx = sx + (0.5 + np.float64(x)) * np.float64(col[j])
        sy = np.float64(-0.0)
        for j, y in enumerate(range(y0, y1)):
            sy = sy + (0.5 + np.float64(y)) * np.float64(row[j])
        lx = sx / np.float64(total)
        ly = sy / np.float64(total)
        lx = lx / W * target_size[0]
        ly = ly / H * target_size[1]
    return cy, cx, lx, ly


def decode_batch(heatmaps, target_size=(224, 224)):
    """[B,K,H,W] float32 -> (idx int64 [B,K,2] = (row, col), xy float64 [B,K,2] = (x, y))."""
    heatmaps = np.asarray(heatmaps, dtype=F32)
    B, K = heatmaps.shape[:2]
    idx = np.zeros((B, K, 2), np.int64)
    xy = np.zeros((B, K, 2), np.float64)
    for b in range(B):
        for k in range(K):
            r, c, x, y = decode_map(heatmaps[b, k], target_size)
            idx[b, k] = (r, c)
            xy[b, k] = (x, y)
    return idx, xy
