"""NumPy emulation of k_forest_predict_ranked's arithmetic on the device image (ranked_image): the
same carry-coded steps, chunk-relative child indices, per-node leaf values and tree-order float64
sums, vectorised over rows.  Lets the image builder be checked against the oracle without a GPU."""
import numpy as np


def slot_ranks(rk, codes_by_attr):
    """int64 [n, n_slots]: what the kernel writes into its rank tile."""
    n = len(next(iter(codes_by_attr.values())))
    out = np.zeros((n, max(int(rk["n_slots"]), 1)), dtype=np.int64)
    off = rk["rank_lut_off"]
    for s, a in enumerate(rk["slot_attr"]):
        code = np.asarray(codes_by_attr[a], dtype=np.int64)
        k = off[s] + code + 1
        ok = (k >= off[s]) & (k < off[s + 1])
        out[:, s] = np.where(ok, rk["rank_lut"][np.clip(k, 0, len(rk["rank_lut"]) - 1)], rk["slot_nan"][s])
    return out


def eval_image(rk, img, baseline, codes_by_attr):
    """-> float64 [n, n_seq] margins exactly as the kernel accumulates them."""
    ranks = slot_ranks(rk, codes_by_attr)
    n = ranks.shape[0]
    rows = np.arange(n)
    depth = int(rk["max_depth"])
    raw = np.tile(np.asarray(baseline, dtype=np.float64), (n, 1))
    hdr = img["tree_hdr"].reshape(-1, 2)
    cto, cs = img["chunk_tree_off"], img["chunk_seq"]
    cn, cl, ch = img["chunk_node_off"], img["chunk_leaf_off"], img["chunk_hdr_off"]
    word = img["word"].astype(np.int64)
    for c in range(len(cs)):
        for j in range(int(cto[c + 1] - cto[c])):
            w = np.full(n, int(hdr[ch[c] + j, 0]), dtype=np.int64)
            bias = int(np.int32(hdr[ch[c] + j, 1]))
            for _ in range(max(depth - 1, 0)):
                w2 = (w + ranks[rows, w >> 24]) & 0xFFFFFFFF
                w = word[cn[c] + ((w2 >> 8) & 0xFFFF)]
            if depth > 0:
                w = (w + ranks[rows, w >> 24]) & 0xFFFFFFFF
            node = (w >> 8) & 0xFFFF
            assert (bias + node).min() >= 0
            raw[:, cs[c]] += img["leaf"][cl[c] + bias + node]
    return raw
