"""Host-side layout conversions between TensorFlow's conventions (NHWC activations, HWIO
conv kernels, [in,out] dense kernels -- what DINCAE/__init__.py:442-513 creates) and the
device layouts of the kernels (planar-4 activations, "WP" conv weights, see
include/dincae_b200.h).  Used at initialisation, for weight import/export and by the
tests; never on the per-step path.
"""
import numpy as np


def planes(c):
    return (c + 3) // 4


def round_up(x, m):
    return (x + m - 1) // m * m


def nhwc_to_p4(x):
    """numpy [B,H,W,C] -> [B,CP,H,W,4] (zero-padded lanes)."""
    B, H, W, C = x.shape
    cp = planes(C)
    out = np.zeros((B, cp, H, W, 4), dtype=x.dtype)
    for c in range(C):
        out[:, c // 4, :, :, c % 4] = x[..., c]
    return out


def p4_to_nhwc(x, C):
    """numpy [B,CP,H,W,4] -> [B,H,W,C]."""
    B, cp, H, W, _ = x.shape
    out = np.empty((B, H, W, C), dtype=x.dtype)
    for c in range(C):
        out[..., c] = x[:, c // 4, :, :, c % 4]
    return out


def cin_map(splits, lanes=4):
    """Packed channel index of every TF input channel for a concatenation of sources with
    ``splits`` channels each (each source padded separately to a multiple of ``lanes``
    channels: 4 = planar-4 fp32, 8 = planar-8 bf16).  Second value: planes of 4 channels."""
    idx = []
    base = 0
    for c in splits:
        idx.extend(range(base, base + c))
        base += round_up(c, lanes)
    return np.array(idx, dtype=np.int64), base // 4


def pack_conv(k_hwio, splits, cout_pad, lanes=4):
    """HWIO [3,3,Cin,Cout] -> WP [9][CinP][cout_pad][4]."""
    kh, kw, cin, cout = k_hwio.shape
    assert (kh, kw) == (3, 3) and sum(splits) == cin and cout <= cout_pad
    idx, cinp = cin_map(splits, lanes)
    wp = np.zeros((9, cinp, cout_pad, 4), dtype=np.float32)
    k = np.asarray(k_hwio, dtype=np.float32).reshape(9, cin, cout)
    wp[:, idx // 4, :cout, idx % 4] = np.transpose(k, (1, 0, 2))  # advanced idx -> [cin,9,cout]
    return wp


def unpack_conv(wp, splits, cout, lanes=4):
    """WP -> HWIO [3,3,Cin,Cout]."""
    idx, cinp = cin_map(splits, lanes)
    assert wp.shape[0] == 9 and wp.shape[1] == cinp
    k = wp[:, idx // 4, :cout, idx % 4]           # [cin, 9, cout]
    return np.ascontiguousarray(np.transpose(k, (1, 0, 2))).reshape(3, 3, len(idx), cout)


def planes8(c):
    return (c + 7) // 8


def nhwc_to_p8(x):
    """numpy [B,H,W,C] -> [B,CP8,H,W,8] (zero-padded lanes; cast to bf16 by the caller)."""
    B, H, W, C = x.shape
    cp = planes8(C)
    out = np.zeros((B, cp, H, W, 8), dtype=x.dtype)
    for c in range(C):
        out[:, c // 8, :, :, c % 8] = x[..., c]
    return out


def p8_to_nhwc(x, C):
    """numpy [B,CP8,H,W,8] -> [B,H,W,C]."""
    B, cp, H, W, _ = x.shape
    out = np.empty((B, H, W, C), dtype=x.dtype)
    for c in range(C):
        out[..., c] = x[:, c // 8, :, :, c % 8]
    return out


def sign_mask8(z_nhwc):
    """Sign mask of the bf16 path: uint8 [B,CP8,H,4*ceil(W/4)], bit j = lane j of the plane > 0."""
    B, H, W, C = z_nhwc.shape
    cp, wq = planes8(C), (W + 3) // 4
    out = np.zeros((B, cp, H, 4 * wq), dtype=np.uint8)
    pos = z_nhwc > 0
    for c in range(C):
        out[:, c // 8, :, :W] |= (pos[..., c].astype(np.uint8) << np.uint8(c % 8))
    return out


def flat_map(H, W, C, cp=None):
    """Planar-flatten index of every NHWC-flatten index of an [H,W,C] tensor stored with
    ``cp`` planes of 4 channels (default: ceil(C/4))."""
    y, x, c = np.meshgrid(np.arange(H), np.arange(W), np.arange(C), indexing="ij")
    cp = planes(C) if cp is None else cp
    return ((((c // 4) * H + y) * W + x) * 4 + c % 4).reshape(-1), cp * H * W * 4


def pack_dense(w, rows_map, rows_n, cols_map, cols_n):
    """TF dense kernel [in,out] -> [rows_n, cols_n] with permuted/padded rows and columns."""
    out = np.zeros((rows_n, cols_n), dtype=np.float32)
    out[np.ix_(rows_map, cols_map)] = w
    return out


def unpack_dense(wp, rows_map, cols_map):
    return np.ascontiguousarray(wp[np.ix_(rows_map, cols_map)])


def sign_mask(z_nhwc):
    """Reference sign mask of a pre-activation [B,H,W,C] -> uint16 [B,CP,H,ceil(W/4)]."""
    B, H, W, C = z_nhwc.shape
    cp, wq = planes(C), (W + 3) // 4
    out = np.zeros((B, cp, H, wq), dtype=np.uint16)
    pos = z_nhwc > 0
    for c in range(C):
        for x in range(W):
            out[:, c // 4, :, x // 4] |= (pos[:, :, x, c].astype(np.uint16)
                                          << np.uint16(4 * (x % 4) + c % 4))
    return out
