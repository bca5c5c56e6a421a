"""Deterministic item-table gradient of the embedding gather (k_tablegrad.cu): grad_table[ids[t]] += dx[t] as a fixed-order segmented
sum over the (id, t)-sorted tokens.  Reference semantics: the scatter-add of tf.gather's gradient (OnDeviceEmbedding,
bert4rec/models/components/networks/bert4rec_encoder.py:103-108).  Integer-valued rows make every summation order exact in fp32, so
the kernel is compared BIT-EXACTLY with torch.index_add_ on the CPU; float rows are compared with a float64 sum and run twice."""
import ctypes as C

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _table_grad(ids, dx, V, base=None):
    from bert4rec_b200 import _lib
    lib = _lib.load()
    T, H = dx.shape
    out = torch.zeros(V, H, device="cuda") if base is None else base.clone()
    ws = torch.empty(lib.b4r_table_grad_workspace_bytes(T, H), dtype=torch.uint8, device="cuda")
    _lib.check(lib.b4r_table_grad(C.c_void_p(ids.data_ptr()), C.c_void_p(dx.data_ptr()), C.c_void_p(out.data_ptr()), T, V, H,
                                  C.c_void_p(ws.data_ptr()), C.c_void_p(torch.cuda.current_stream().cuda_stream)))
    torch.cuda.synchronize()
    return out


def _ids(kind, T, V, rng):
    if kind == "uniform":
        return rng.randint(0, V, size=T)
    if kind == "zipf":                       # the bench's distribution: a few very long runs (MASK / popular items) and a long tail
        x = (rng.zipf(1.1, size=T) - 1) % (V - 3) + 3
        x[rng.rand(T) < 0.2] = 2
        return x
    if kind == "one_item":                   # a single run over every chunk
        return np.full(T, V - 1)
    if kind == "chunk_aligned":              # runs that begin and end exactly on chunk boundaries, and runs of one token between them
        x = np.repeat(np.arange(T // 64 + 1) * 2 % V, 64)[:T]
        x[::193] = V - 1
        return x
    if kind == "out_of_range":               # ids outside [0, V) are clamped like the gather clamps them
        x = rng.randint(-5, V + 5, size=T)
        return x
    raise KeyError(kind)


@pytest.mark.parametrize("kind", ["uniform", "zipf", "one_item", "chunk_aligned", "out_of_range"])
@pytest.mark.parametrize("T,V,H", [(1, 7, 64), (63, 100, 64), (64, 5, 128), (4097, 300, 256), (12800, 12004, 64), (51200, 70000, 128),
                                   (204800, 13047, 256), (40000, 1000003, 256), (600000, 5003, 64)])
def test_table_grad_is_exact_on_integer_rows(kind, T, V, H):
    rng = np.random.RandomState(T + V + H)
    ids_np = _ids(kind, T, V, rng).astype(np.int64)
    ids = torch.from_numpy(ids_np).cuda()
    dx = torch.from_numpy(rng.randint(-8, 9, size=(T, H)).astype(np.float32)).cuda()
    base = torch.from_numpy(rng.randint(-3, 4, size=(V, H)).astype(np.float32)).cuda()
    got = _table_grad(ids, dx, V, base)
    want = base.clone().index_add_(0, ids.clamp(0, V - 1), dx)          # exact for small integers whatever the order
    assert torch.equal(got, want)


@pytest.mark.parametrize("kind", ["zipf", "one_item"])
def test_table_grad_float_rows_fixed_order(kind):
    T, V, H = 204800, 13047, 256
    rng = np.random.RandomState(5)
    ids = torch.from_numpy(_ids(kind, T, V, rng).astype(np.int64)).cuda()
    dx = torch.randn(T, H, device="cuda") * torch.rand(T, 1, device="cuda") * 1e-2
    a = _table_grad(ids, dx, V)
    b = _table_grad(ids, dx, V)
    assert torch.equal(a, b)                                              # bit-reproducible
    want = torch.zeros(V, H, dtype=torch.float64, device="cuda").index_add_(0, ids, dx.double())
    err = (a.double() - want).abs().max().item()
    scale = want.abs().max().item()
    assert err < 1e-5 * scale + 1e-7, (err, scale)
