"""The hand-written radix passes of the binning stage (csrc/chs_sort.cuh) on adversarial keys, through the C ABI
(chs_radix_sort_pairs): stability, segments that never mix, skewed digits, sizes around the 4096-item tile."""
import ctypes

import pytest
import torch

from casualhdrsplat_b200 import _lib

pytestmark = pytest.mark.gpu


def _sort(keys: torch.Tensor, n_seg: int, seg_len: int, bits: int):
    L = _lib.lib()
    dev = keys.device
    n = n_seg * seg_len
    ko = torch.empty(n, dtype=torch.int32, device=dev)
    vo = torch.empty(n, dtype=torch.int32, device=dev)
    tiles = n_seg * ((seg_len + 4095) // 4096)
    ws = torch.empty(8 * n + 1024 * (tiles + n_seg) + 4096, dtype=torch.uint8, device=dev)
    _lib.check(L.chs_radix_sort_pairs(_lib.ptr(keys), n_seg, seg_len, bits, _lib.ptr(ko), _lib.ptr(vo), _lib.ptr(ws), ws.numel(),
                                      ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)), "chs_radix_sort_pairs")
    torch.cuda.synchronize()
    return ko, vo


@pytest.mark.parametrize("n_seg,seg_len", [(1, 1), (1, 31), (1, 4096), (1, 4097), (3, 5000), (8, 100_003), (5, 12_288), (130, 77)])
@pytest.mark.parametrize("dist", ["uniform", "few", "constant", "sorted", "depth"])
def test_segmented_radix_sort_is_stable_and_exact(n_seg, seg_len, dist):
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(n_seg * 1000 + seg_len)
    n = n_seg * seg_len
    if dist == "uniform":
        k = torch.randint(0, 2**31 - 1, (n,), generator=g, dtype=torch.int64) * 2 + torch.randint(0, 2, (n,), generator=g)
    elif dist == "few":  # three distinct values: every tile is dominated by one digit in every pass
        k = torch.tensor([7, 0xFFFFFFFF, 0x40400000])[torch.randint(0, 3, (n,), generator=g)]
    elif dist == "constant":
        k = torch.full((n,), 0x3F800000, dtype=torch.int64)
    elif dist == "sorted":
        k = torch.arange(n, dtype=torch.int64) * 977 % (2**32)
        k = torch.sort(k).values
    else:  # positive floats with many ties, as camera-space depths with culled entries
        z = (2.0 + 10.0 * torch.rand(n, generator=g)).float()
        z[torch.rand(n, generator=g) < 0.2] = 4.0
        k = z.view(torch.int32).to(torch.int64)
        k[torch.rand(n, generator=g) < 0.1] = 0xFFFFFFFF
    keys = (k & 0xFFFFFFFF).to(torch.int64)
    as_i32 = torch.where(keys >= 2**31, keys - 2**32, keys).to(torch.int32).to(dev)
    ko, vo = _sort(as_i32, n_seg, seg_len, 32)
    want_k, want_v = [], []
    for s in range(n_seg):
        seg = keys[s * seg_len:(s + 1) * seg_len]
        sk, perm = torch.sort(seg, stable=True)
        want_k.append(sk)
        want_v.append(perm + s * seg_len)
    want_k, want_v = torch.cat(want_k), torch.cat(want_v)
    assert torch.equal(ko.cpu().to(torch.int64) & 0xFFFFFFFF, want_k)
    assert torch.equal(vo.cpu().to(torch.int64), want_v)


def test_partial_key_width_sorts_only_the_low_bits():
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(3)
    n = 50_000
    keys = torch.randint(0, 2**20, (n,), generator=g, dtype=torch.int64)
    ko, vo = _sort(keys.to(torch.int32).to(dev), 1, n, 13)  # two passes: bits 0..15 take part, higher bits ride along
    sk, perm = torch.sort(keys & 0xFFFF, stable=True)
    assert torch.equal(vo.cpu().to(torch.int64), perm)
    assert torch.equal(ko.cpu().to(torch.int64), keys[perm])
