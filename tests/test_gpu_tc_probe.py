"""tcgen05 building blocks (noc_tc_probe): UMMA descriptors in both majors, TMEM round trip, against torch matmul."""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("N,K,mn,rt", [(128, 128, 0, 0), (128, 128, 1, 0), (128, 16, 0, 0), (16, 128, 1, 0), (16, 16, 0, 0),
                                       (64, 64, 1, 1), (128, 128, 0, 1)])
def test_umma_probe_matches_matmul(N, K, mn, rt):
    import neuraloc_b200 as nb
    lib = nb._cabi.lib()
    g = torch.Generator(device="cuda").manual_seed(N * 1000 + K + mn)
    A = torch.randn(128, K, generator=g, device="cuda").bfloat16()
    if mn:
        B = torch.randn(K, N, generator=g, device="cuda").bfloat16()       # D = A @ B
        ref = A.float() @ B.float()
    else:
        B = torch.randn(N, K, generator=g, device="cuda").bfloat16()       # D = A @ B^T
        ref = A.float() @ B.float().t()
    D = torch.full((128, N), float("nan"), device="cuda")
    nb._cabi.check(lib.noc_tc_probe(A.data_ptr(), B.data_ptr(), D.data_ptr(), N, K, mn, rt, None))
    torch.cuda.synchronize()
    err = (D - ref).abs().max().item()
    assert err <= 1e-3 * max(1.0, ref.abs().max().item()), (N, K, mn, rt, err)
