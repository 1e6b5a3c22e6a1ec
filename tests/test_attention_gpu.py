"""GPU: the non-local block on the library's own kernels (attention.py, csrc/attention_ops.cu) against the composite
torch.bmm / softmax / max_pool2d formulation of the reference (u_net_2d_discriminator.py:332-381), which the module
still runs under _mode.higher_order_gradients(), and against the CPU oracle."""
import pytest
import torch
import torch.nn.functional as F

from tests.conftest import rel_err

pytestmark = pytest.mark.gpu


def dev():
    return torch.device("cuda:0")


def cl(t):
    return t.contiguous(memory_format=torch.channels_last)


def test_split_pool_and_softmax_passes(built_library):
    from multi_stylegan_b200 import _C
    torch.manual_seed(0)
    B, H, W, cq, cv = 3, 11, 14, 8, 20          # odd height: the last row is not pooled
    qkv = cl(torch.randn(B, 2 * cq + cv, H, W, device=dev()))
    qkv[:, cq:, 0:2, 0:2] = 1.5                 # a tie in the first window: the first position wins
    theta, phi, g, idx = _C.nl_split_pool(qkv, cq, cv)
    assert torch.equal(theta, qkv[:, :cq])
    want, widx = F.max_pool2d(qkv[:, cq:], 2, 2, return_indices=True)
    assert torch.equal(phi, want[:, :cq]) and torch.equal(g, want[:, cq:])
    dth, dph, dg = torch.randn_like(theta), torch.randn_like(phi), torch.randn_like(g)
    dq = _C.nl_merge_unpool(dth, dph, dg, idx)
    ref = torch.cat([dth, F.max_unpool2d(torch.cat([dph, dg], 1), widx, 2, 2, output_size=(H, W))], 1)
    assert torch.equal(dq, ref)
    for n in (100, 1024, 2048 + 4):
        x = torch.randn(37, n, device=dev()) * 3
        p = _C.softmax_rows_(x.clone(), n)
        want = torch.softmax(x.double(), -1)
        assert (p.double() - want).abs().max() < 1e-6
        gp = torch.randn_like(x)
        ds = _C.softmax_rows_bwd_(gp.clone(), p, n)
        wd = want * (gp.double() - (want * gp.double()).sum(-1, keepdim=True))
        assert (ds.double() - wd).abs().max() < 1e-5


@pytest.mark.parametrize("case", [(2, 64, 96, 20, 20, False), (2, 32, 64, 16, 24, True), (1, 384, 384, 64, 64, True)],
                         ids=["enc-20x20", "dec-cat-16x24", "default-decoder-64x64"])
def test_non_local_block_matches_composite(built_library, case):
    from multi_stylegan_b200 import _mode
    import multi_stylegan_b200.u_net_2d_discriminator as D_mod
    B, cin, cout, H, W, two = case
    torch.manual_seed(H + W)
    blk = D_mod.NonLocalBlock(cin * (2 if two else 1), cout).to(dev())
    with torch.no_grad():
        blk.gamma.fill_(0.7)
    x = cl(torch.randn(B, cin, H, W, device=dev())).requires_grad_(True)
    x2 = cl(torch.randn(B, cin, H, W, device=dev())).requires_grad_(True) if two else None
    gout = cl(torch.randn(B, cout, H, W, device=dev()))

    def run(fused):
        for t in [x, x2] + list(blk.parameters()):
            if t is not None:
                t.grad = None
        if fused:
            out = blk(x, x2)
            out.backward(gout)
        else:
            with _mode.higher_order_gradients():
                out = blk(x, x2)
            out.backward(gout)
        grads = {"x": x.grad.clone()}
        if two:
            grads["x2"] = x2.grad.clone()
        grads.update({n: p.grad.clone() for n, p in blk.named_parameters()})
        return out.detach().clone(), grads

    got, g1 = run(True)
    want, g2 = run(False)
    assert rel_err(got, want) < 1e-2, rel_err(got, want)
    assert set(g1) == set(g2)
    for n in g1:
        # TF32 on both sides (the composite runs its matmuls in TF32 like PyTorch 1.8.1, the fused one on tcgen05 TF32)
        assert rel_err(g1[n], g2[n]) < 2e-2, (n, rel_err(g1[n], g2[n]))


def test_non_local_block_matches_oracle(built_library):
    from oracle import model as omodel
    import multi_stylegan_b200.u_net_2d_discriminator as D_mod
    torch.manual_seed(3)
    blk = D_mod.NonLocalBlock(64, 96)
    with torch.no_grad():
        blk.gamma.fill_(0.5)
    sd = {"b." + k: v.detach().clone() for k, v in blk.state_dict().items()}
    x = torch.randn(2, 64, 24, 24)
    want = omodel.non_local_block(sd, "b", x)
    got = blk.to(dev())(cl(x.to(dev())))
    assert rel_err(got, want) < 1e-2, rel_err(got, want)


@pytest.mark.parametrize("case", [(8, 12, 5, 7, 1), (16, 384, 32, 32, 2), (6, 20, 4, 4, 3), (4, 768, 16, 16, 1)],
                         ids=lambda c: "B%d-C%d-%dx%d-groups%d" % c)
def test_minibatch_stddev_kernels_match_tensor_op_formulation(built_library, case):
    """msg_mbstd_forward / _backward vs the module's tensor-op formulation (u_net_2d_discriminator.py:189-217), which it still
    runs under higher_order_gradients()."""
    from multi_stylegan_b200 import _mode
    import multi_stylegan_b200.u_net_2d_discriminator as D_mod
    B, C, H, W, groups = case
    torch.manual_seed(B + C)
    m = D_mod.MinibatchStdDev()
    m.groups = groups
    x = cl(torch.randn(B, C, H, W, device=dev()) * 2).requires_grad_(True)
    x.data[:, 0, 0, 0] = 1.25                      # one position with zero variance: the clamp branch
    gout = cl(torch.randn(B, C + 1, H, W, device=dev()))
    got = m(x)
    got.backward(gout)
    g1 = x.grad.clone()
    x.grad = None
    with _mode.higher_order_gradients():
        want = m(x)
    want.backward(gout)
    assert got.shape == want.shape
    assert rel_err(got, want) < 1e-6, rel_err(got, want)
    assert rel_err(g1, x.grad) < 1e-5, rel_err(g1, x.grad)
