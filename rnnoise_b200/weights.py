"""Weight tooling for the batched engine: read/write the reference's binary weight blob and export a
PyTorch RNNoise checkpoint straight to a blob.

The reference gets from a checkpoint to `weights_blob.bin` in three steps: the exporter
`torch/rnnoise/dump_rnnoise_weights.py --quantize` prints ~30 MB of C arrays
(`wexchange/c_export/common.py`), a C compiler parses them, and `src/write_weights.c:46-69`
(`dump_weights_blob`) serialises the arrays as "DNNw" records.  `export_state_dict` below produces the
same bytes without the C detour; `tests/test_weights.py` pins it byte-for-byte against blobs made by
the reference pipeline (`oracle/make_models.py`).

Blob format (`nnet.h:55-62`, `write_weights.c:51-66`): a sequence of records, each a 64-byte head
  char head[4] = "DNNw"; int version = 0; int type; int size; int block_size; char name[44];
followed by `block_size` payload bytes (`size` rounded up to a multiple of 64, zero padded).
Types (`nnet.h:50-53`): 0 float, 1 int, 2 qweight, 3 int8.  Machine (little) endian.

What the exporter has to reproduce (file:line in /root/reference/torch/weight-exchange/wexchange):
  * layer walk order conv1, conv2, gru1..3, dense_out, vad_dense; conv1 and the two heads stay float
    (`dump_rnnoise_weights.py:15,62-76`), everything else is int8 with per-output scales;
  * matrices are stored input-major, `W[j * n_out + i]` (`c_export/common.py:250-252,266-271`);
    conv taps flatten as (tap, in) (`:263-266`); GRU gates are reordered r,z,n -> z,r,n (`:343-350`);
  * per-output scale = max(max|w| / 127, max|w[2k] + w[2k+1]| / 129) (`:164-178`), stored / 127 (`:238`);
    q = round(w / scale) (`:122-128`); subias = bias - sum_j q[j] * scale (`:235`);
  * dense int8 matrices are tiled [out/8][in/4][8][4] (`:57-59`); sparse ones list, per block of 8
    outputs, the number of non-zero 4-input blocks, then their first input index, with the weights of
    each kept block in [8][4] order (`:131-160`); the recurrent diagonal is split off first (`:103-120`).
Array dtypes and memory orders below follow the exporter's numpy expressions where they decide the
rounding of `subias` (a float64 sum whose order depends on the operand layout).
"""
import struct

import numpy as np

BLOCK = 64
TYPE_FLOAT, TYPE_INT, TYPE_QWEIGHT, TYPE_INT8 = 0, 1, 2, 3
_DTYPES = {TYPE_FLOAT: np.float32, TYPE_INT: np.int32, TYPE_QWEIGHT: np.int8, TYPE_INT8: np.int8}
_HEAD = struct.Struct("<4siiii44s")


# ------------------------------------------------------------------------------------------------
# blob container
# ------------------------------------------------------------------------------------------------
def read_blob(data):
    """bytes -> [(name, type, ndarray)] in file order; raises ValueError on a malformed record
    (the checks of parse_lpcnet_weights.c:39-43)."""
    if isinstance(data, str):
        with open(data, "rb") as f:
            data = f.read()
    out, off = [], 0
    while off < len(data):
        if len(data) - off < BLOCK:
            raise ValueError("truncated record head")
        head, version, typ, size, block, name = _HEAD.unpack_from(data, off)
        if head != b"DNNw" or version != 0 or size < 0 or block < size or block % BLOCK or block > len(data) - off - BLOCK:
            raise ValueError(f"bad record at offset {off}")
        if b"\0" not in name:
            raise ValueError("unterminated array name")
        if typ not in _DTYPES:
            raise ValueError(f"unknown array type {typ}")
        payload = data[off + BLOCK: off + BLOCK + size]
        out.append((name.split(b"\0")[0].decode(), typ, np.frombuffer(payload, dtype=_DTYPES[typ]).copy()))
        off += BLOCK + block
    return out


def write_blob(records, path=None):
    """[(name, type, array)] -> bytes (and to `path` if given): the `dump_weights_blob` step."""
    chunks = []
    for name, typ, arr in records:
        raw = np.ascontiguousarray(arr, dtype=_DTYPES[typ]).tobytes()
        if len(name) >= 43:
            raise ValueError(f"array name too long: {name}")
        block = (len(raw) + BLOCK - 1) // BLOCK * BLOCK
        chunks.append(_HEAD.pack(b"DNNw", 0, typ, len(raw), block, name.encode()))
        chunks.append(raw + bytes(block - len(raw)))
    blob = b"".join(chunks)
    if path is not None:
        with open(path, "wb") as f:
            f.write(blob)
    return blob


# ------------------------------------------------------------------------------------------------
# checkpoint -> records
# ------------------------------------------------------------------------------------------------
def _f32(x):
    return np.asarray(x).astype(np.float32)   # float64 -> float32 rounds to nearest even, as the C compiler does


def _output_scales(w):
    """w: [n_in, n_out] float32 -> per-output quantisation step (float32)."""
    n_in, n_out = w.shape
    if n_in % 4 or n_out % 8:
        raise ValueError("int8 layers need n_in % 4 == 0 and n_out % 8 == 0")
    peak = np.max(np.abs(w), axis=0)
    pair = np.max(np.abs(w[0:n_in:2] + w[1:n_in:2]), axis=0)   # u8 x s8 pairs must not overflow s16 (maddubs)
    return np.maximum(peak / 127, pair / 129)


def _quantise(w, scale):
    q = np.round(w / (scale + 1e-30)).astype("int")
    if q.max() > 127 or q.min() <= -128:
        raise ValueError("weight does not fit int8 at the computed scale")
    return q


def _int8_layer(name, w, bias, sparse, diagonal):
    """Records of one quantised LinearLayer.  w: [n_in, n_out] float32 in the exporter's memory order."""
    n_in, n_out = w.shape
    scale = _output_scales(w)
    rec = []
    if sparse:
        a = w
        if diagonal:   # recurrent matrices: the per-gate diagonals live in a float vector of their own
            a = w.copy()   # (C order, as the exporter's copy: decides the summation order of subias)
            gates = n_out // n_in
            d = np.concatenate([np.diag(a[:, g * n_in:(g + 1) * n_in]).copy() for g in range(gates)])
            for g in range(gates):
                a[:, g * n_in:(g + 1) * n_in] -= np.diag(d[g * n_in:(g + 1) * n_in])
            rec.append((name + "_weights_diag", TYPE_FLOAT, d))
        q = _quantise(a, scale)
        # [n_in/4, 4, n_out/8, 8] -> per (out block, in block): keep when the FLOAT block is non-zero
        fb = np.abs(a).reshape(n_in // 4, 4, n_out // 8, 8).sum(axis=(1, 3)) > 1e-10      # [in blk, out blk]
        qb = q.reshape(n_in // 4, 4, n_out // 8, 8).transpose(2, 0, 3, 1)                 # [out blk, in blk, 8, 4]
        idx, vals = [], []
        for ob in range(n_out // 8):
            kept = np.nonzero(fb[:, ob])[0]
            idx.append(len(kept))
            idx.extend((4 * kept).tolist())
            vals.append(qb[ob, kept].reshape(-1))
        rec.append((name + "_weights_int8", TYPE_INT8, np.concatenate(vals) if vals else np.zeros(0, np.int8)))
        rec.append((name + "_weights_idx", TYPE_INT, np.asarray(idx, dtype=np.int32)))
    else:
        q = _quantise(w, scale)
        rec.append((name + "_weights_int8", TYPE_INT8, q.reshape(n_in // 4, 4, n_out // 8, 8).transpose(2, 0, 3, 1).reshape(-1)))
    b = np.zeros(n_out) if bias is None else bias
    rec.append((name + "_subias", TYPE_FLOAT, _f32(b - np.sum(q * scale, axis=0))))
    rec.append((name + "_scale", TYPE_FLOAT, _f32(scale / 127 * np.ones(n_out))))
    if bias is not None:
        rec.append((name + "_bias", TYPE_FLOAT, bias))
    return rec


def _float_layer(name, w, bias):
    return [(name + "_weights_float", TYPE_FLOAT, np.reshape(w, -1)), (name + "_bias", TYPE_FLOAT, bias)]


def _np(sd, key):
    v = sd[key]
    if hasattr(v, "detach"):
        v = v.detach().cpu().numpy()
    return np.array(v, dtype=np.float32, order="C")


def export_state_dict(sd, path=None):
    """RNNoise state dict (torch tensors or numpy arrays; keys conv1.weight ... vad_dense.bias as in
    torch/rnnoise/rnnoise.py:66-72) -> blob bytes, int8-quantised like `--quantize`."""
    rec = []
    # conv: torch (out, in, tap) -> rows (tap, in), columns out
    for name, quant in (("conv1", False), ("conv2", True)):
        w = np.transpose(_np(sd, name + ".weight"), (2, 1, 0))
        w = np.reshape(w, (-1, w.shape[-1]))
        b = _np(sd, name + ".bias")
        rec += _int8_layer(name, w, b, sparse=False, diagonal=False) if quant else _float_layer(name, w, b)
    for name in ("gru1", "gru2", "gru3"):
        parts = [_np(sd, f"{name}.{k}_l0") for k in ("weight_ih", "weight_hh", "bias_ih", "bias_hh")]
        n = parts[0].shape[0] // 3
        for x in parts:   # gate order r,z,n -> z,r,n
            x[:2 * n] = np.concatenate([x[n:2 * n], x[:n]])
        w_ih, w_hh, b_ih, b_hh = parts
        rec += _int8_layer(name + "_input", w_ih.transpose(), b_ih, sparse=True, diagonal=False)
        rec += _int8_layer(name + "_recurrent", w_hh.transpose(), b_hh, sparse=True, diagonal=True)
    for name in ("dense_out", "vad_dense"):
        rec += _float_layer(name, _np(sd, name + ".weight").transpose(), _np(sd, name + ".bias"))
    return write_blob(rec, path)


def export_checkpoint(ckpt_path, out_path):
    """A training checkpoint as train_rnnoise.py saves it ({'state_dict': ...}; .pth) or an .npz of
    the same arrays -> blob file.  torch is imported only for .pth files."""
    if ckpt_path.endswith(".npz"):
        sd = dict(np.load(ckpt_path))
    else:
        import torch
        ck = torch.load(ckpt_path, map_location="cpu")
        sd = ck.get("state_dict", ck)
    return export_state_dict(sd, out_path)


def describe(blob):
    """Model dimensions inferred from the array sizes, as the C loader does (model_blob.c)."""
    arrays = {n: a for n, _, a in read_blob(blob)}
    cond = arrays["conv1_bias"].size
    gru = arrays["conv2_bias"].size
    nnz = {n[:-len("_weights_int8")]: a.size for n, a in arrays.items() if n.endswith("_weights_int8")}
    return {"cond": cond, "gru": gru, "int8_weights": nnz, "bytes": sum(a.nbytes for a in arrays.values())}


if __name__ == "__main__":
    import sys
    if len(sys.argv) == 3:
        export_checkpoint(sys.argv[1], sys.argv[2])
        print(describe(sys.argv[2]))
    elif len(sys.argv) == 2:
        print(describe(sys.argv[1]))
    else:
        sys.exit("usage: python -m rnnoise_b200.weights <checkpoint.pth|.npz> <weights_blob.bin>   |   <weights_blob.bin>")
