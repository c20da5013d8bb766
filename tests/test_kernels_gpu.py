"""Kernel-level parity of the tcgen05 conv / wgrad kernels against plain PyTorch fp32 (GPU only).

Inputs and weights are pre-rounded to fp16 (the operand type of the tcgen05 kernels: same 10-bit mantissa as
tf32) so that the only difference between the tensor-core kernel and the fp32 torch reference is accumulation
order and the tf32/fp16 rounding of the STORED result (tolerances written next to each assert).
"""
import ctypes

import pytest
import torch
import torch.nn.functional as F

from probnmn_clevr_b200 import _lib as L
from probnmn_clevr_b200.planes import (GUARD_FLOATS, fmt_for_dilation, from_planes, map_from_slots, round_tf32,
                                       to_half_planes, to_planes)

pytestmark = pytest.mark.gpu


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr())


def _pack(W, n_kb, ntaps, flip, k_off, n_off, k_stride, n_stride, tap_stride):
    """run the library's pack kernel on one weight tensor -> packed float tensor"""
    lib = L.lib()
    t = L.PackTask(0, 0, 0, n_kb, ntaps, flip, k_off, n_off, k_stride, n_stride, tap_stride, 0)
    packed = torch.zeros(n_kb * ntaps * 2048, device="cuda", dtype=torch.float16)
    L.check(lib.pnmn_debug_pack(ctypes.byref(t), 1, n_kb * ntaps, _ptr(W), _ptr(packed), _stream()))
    return packed


def _arena(planes_list):
    """concatenate plane tensors into one zero-guarded arena; return (arena, [float offsets])"""
    total = 2 * GUARD_FLOATS + sum(p.numel() for p in planes_list)
    arena = torch.zeros(total, device="cuda")
    offs, o = [], GUARD_FLOATS
    for p in planes_list:
        arena[o:o + p.numel()] = p.reshape(-1)
        offs.append(o)
        o += p.numel()
    return arena, offs


def _half_arena(hplanes_list):
    total = 2 * GUARD_FLOATS * 2 + sum(p.numel() for p in hplanes_list)
    arena = torch.zeros(total, device="cuda", dtype=torch.float16)
    offs, o = [], GUARD_FLOATS * 2
    for p in hplanes_list:
        arena[o:o + p.numel()] = p.reshape(-1)
        offs.append(o)
        o += p.numel()
    return arena, offs


def _conv_cfg(n_kb, kb_per_in, ntaps, dil, fin, fout, faux, flags):
    lead = ((dil * fin[0] + dil + 7) // 8) * 8 if ntaps == 9 else 8
    return L.ConvCfg(n_kb, kb_per_in, ntaps, dil, fin[0], fin[1], fout[0], fout[1], faux[0], faux[1], flags, lead)


def _skip_without_twins(impl):
    """impl = 1 selects the CUDA-core twin of a tensor-core kernel; the release library is built without the twins
    (probnmn_clevr_b200/csrc/Makefile: BRINGUP=1 adds them) -- the tensor-core kernels are then checked against torch alone"""
    if impl and not L.lib().pnmn_has_bringup_kernels():
        pytest.skip("library built without the CUDA-core bring-up kernels (make BRINGUP=1)")


def _run_conv(cfg, variant, impl, ins, w_packed, bias=None, aux=None, w3=None, b3=None, out_init=None):
    _skip_without_twins(impl)
    """ins: list over samples of list over inputs of (C,14,14) tensors. returns (outs nchw, maps)"""
    lib = L.lib()
    ns = len(ins)
    S_in, S_out = cfg.S_in, cfg.S_out
    in_planes = [to_half_planes(x, S_in) for smp in ins for x in smp]
    arena_in, offs_in = _half_arena(in_planes)
    out_planes = [to_planes(out_init[s], S_out) if out_init is not None else torch.zeros(32, S_out * S_out, 4, device="cuda")
                  for s in range(ns)]
    arena_out, offs_out = _arena(out_planes)
    arena_aux, offs_aux = (None, None)
    if aux is not None:
        arena_aux, offs_aux = _arena([to_planes(a, cfg.S_aux) for a in aux])
    maps = torch.zeros(ns, 256, device="cuda")
    t = L.ConvTask()
    n_in = len(ins[0])
    for s in range(ns):
        for i in range(n_in):
            t.in_[i][s] = arena_in.data_ptr() + 2 * offs_in[s * n_in + i]
        t.out[s] = arena_out.data_ptr() + 4 * offs_out[s]
        if aux is not None:
            t.aux[s] = arena_aux.data_ptr() + 4 * offs_aux[s]
        t.map_out[s] = maps.data_ptr() + 4 * 256 * s
    t.w = w_packed.data_ptr()
    t.bias = bias.data_ptr() if bias is not None else None
    t.w3 = w3.data_ptr() if w3 is not None else None
    t.b3 = b3.data_ptr() if b3 is not None else None
    t.cfg = 0
    t.n_samp = ns
    t.mt0, t.n_mt = 0, (3 if variant == 1 else 2)
    L.check(lib.pnmn_debug_launch_conv(ctypes.byref(t), 1, ctypes.byref(cfg), 1, variant, impl, _stream()))
    torch.cuda.synchronize()
    outs = []
    for s in range(ns):
        n = 32 * S_out * S_out * 4
        outs.append(from_planes(arena_out[offs_out[s]:offs_out[s] + n].view(32, S_out * S_out, 4), S_out))
    return outs, [map_from_slots(maps[s]) for s in range(ns)]


def _relerr(a, b):
    return (a - b).abs().max().item() / (b.abs().max().item() + 1e-30)


def _mk(shape, gen, scale=1.0):
    return (torch.randn(shape, generator=gen, device="cuda") * scale).half().float()


@pytest.mark.parametrize("impl", [1, 0], ids=["simt", "tcgen05"])
@pytest.mark.parametrize("dil,ns", [(1, 2), (1, 1), (2, 2), (4, 2), (8, 1)])
def test_conv3x3_forward(impl, dil, ns):
    g = torch.Generator(device="cuda").manual_seed(dil * 10 + ns)
    W = _mk((128, 128, 3, 3), g, 0.05)
    b = torch.randn(128, generator=g, device="cuda")
    xs = [_mk((128, 14, 14), g) for _ in range(ns)]
    fin = fmt_for_dilation(dil)
    fout = fmt_for_dilation({1: 2, 2: 4, 4: 8, 8: 1}[dil])  # Relate chain: next conv's input format
    cfg = _conv_cfg(8, 8, 9, dil, fin, fout, fout, L.F_BIAS | L.F_RELU | L.F_STORE)
    wp = _pack(W, 8, 9, 0, 0, 0, 9, 128 * 9, 1)
    outs, _ = _run_conv(cfg, 1 if fin[0] == 22 else 0, impl, [[x] for x in xs], wp, bias=b)
    for x, o in zip(xs, outs):
        ref = F.relu(F.conv2d(x[None], W, b, padding=dil, dilation=dil))[0]
        err = _relerr(o, round_tf32(ref))
        print(f"conv3x3 impl={impl} dil={dil} ns={ns}: rel err {err:.3e}")
        assert err < 1e-3  # output is stored tf32-rounded: 2^-11 = 4.9e-4 worst case


@pytest.mark.parametrize("impl", [1, 0], ids=["simt", "tcgen05"])
def test_conv_dotsig_head(impl):
    g = torch.Generator(device="cuda").manual_seed(7)
    W = _mk((128, 128, 3, 3), g, 0.05)
    b = torch.randn(128, generator=g, device="cuda")
    w3 = torch.randn(128, generator=g, device="cuda") * 0.05
    b3 = torch.randn(1, generator=g, device="cuda")
    xs = [_mk((128, 14, 14), g) for _ in range(2)]
    f16 = fmt_for_dilation(1)
    cfg = _conv_cfg(8, 8, 9, 1, f16, f16, f16, L.F_BIAS | L.F_RELU | L.F_STORE | L.F_DOTSIG)
    wp = _pack(W, 8, 9, 0, 0, 0, 9, 128 * 9, 1)
    _, maps = _run_conv(cfg, 0, impl, [[x] for x in xs], wp, bias=b, w3=w3, b3=b3)
    for x, m in zip(xs, maps):
        y = F.relu(F.conv2d(x[None], W, b, padding=1))
        ref = torch.sigmoid(F.conv2d(y, w3.view(1, 128, 1, 1), b3))[0, 0]
        err = (m - ref).abs().max().item()
        print(f"dotsig impl={impl}: abs err {err:.3e}")
        assert err < 1e-5  # fp32 epilogue on exact tf32 products


@pytest.mark.parametrize("impl", [1, 0], ids=["simt", "tcgen05"])
@pytest.mark.parametrize("mode", ["mask", "accum"])
@pytest.mark.parametrize("dil", [1, 8])
def test_conv_dgrad_mask_accum(impl, dil, mode):
    """dgrad = same kernel on the transposed/flipped pack; the epilogue either masks with Y>0 (ReLU backward)
    or accumulates into the existing gradient (the scheduler never asks for both at once)"""
    g = torch.Generator(device="cuda").manual_seed(11 + dil)
    W = _mk((128, 128, 3, 3), g, 0.05)
    ns = 1 if dil == 8 else 2
    dz = [_mk((128, 14, 14), g) for _ in range(ns)]
    yprev = [_mk((128, 14, 14), g) for _ in range(ns)]
    old = [_mk((128, 14, 14), g) for _ in range(ns)]
    fin = fmt_for_dilation(dil)
    f16 = fmt_for_dilation(1)
    cfg = _conv_cfg(8, 8, 9, dil, fin, f16, fin, L.F_STORE | (L.F_MASK if mode == "mask" else L.F_ACCUM))
    wp = _pack(W, 8, 9, 1, 0, 0, 128 * 9, 9, 1)
    outs, _ = _run_conv(cfg, 1 if fin[0] == 22 else 0, impl, [[d] for d in dz], wp, aux=yprev, out_init=old)
    for d, y, o0, o in zip(dz, yprev, old, outs):
        gi = torch.nn.grad.conv2d_input((1, 128, 14, 14), W, d[None], padding=dil, dilation=dil)[0]
        ref = torch.where(y > 0, gi, torch.zeros_like(gi)) if mode == "mask" else gi + o0
        err = _relerr(o, ref)
        print(f"dgrad impl={impl} dil={dil}: rel err {err:.3e}")
        assert err < 1e-3  # tf32-rounded store


@pytest.mark.parametrize("impl", [1, 0], ids=["simt", "tcgen05"])
def test_conv_projection_two_inputs(impl):
    g = torch.Generator(device="cuda").manual_seed(5)
    W = _mk((128, 256, 1, 1), g, 0.05)
    b = torch.randn(128, generator=g, device="cuda")
    ins = [[_mk((128, 14, 14), g), _mk((128, 14, 14), g)] for _ in range(2)]
    f16 = fmt_for_dilation(1)
    cfg = _conv_cfg(16, 8, 1, 1, f16, f16, f16, L.F_BIAS | L.F_RELU | L.F_STORE)
    wp = _pack(W, 16, 1, 0, 0, 0, 1, 256, 0)
    outs, _ = _run_conv(cfg, 0, impl, ins, wp, bias=b)
    for (a, c), o in zip(ins, outs):
        ref = F.relu(F.conv2d(torch.cat([a, c])[None], W, b))[0]
        err = _relerr(o, ref)
        print(f"projection impl={impl}: rel err {err:.3e}")
        assert err < 1e-3


@pytest.mark.parametrize("impl", [1, 0], ids=["simt", "tcgen05"])
def test_conv_stem_1024(impl):
    g = torch.Generator(device="cuda").manual_seed(3)
    W = _mk((128, 1024, 3, 3), g, 0.02)
    b = torch.randn(128, generator=g, device="cuda")
    xs = [_mk((1024, 14, 14), g).relu() for _ in range(2)]
    f16 = fmt_for_dilation(1)
    cfg = _conv_cfg(64, 64, 9, 1, f16, f16, f16, L.F_BIAS | L.F_RELU | L.F_STORE)
    wp = _pack(W, 64, 9, 0, 0, 0, 9, 1024 * 9, 1)
    outs, _ = _run_conv(cfg, 0, impl, [[x] for x in xs], wp, bias=b)
    for x, o in zip(xs, outs):
        ref = F.relu(F.conv2d(x[None], W, b, padding=1))[0]
        err = _relerr(o, ref)
        print(f"stem impl={impl}: rel err {err:.3e}")
        assert err < 1e-3


def _run_wgrad(impl, dzs, xs, dil, ksize, cin_total, cin0, scale=1.0):
    _skip_without_twins(impl)
    """dzs / xs: fp16-representable tensors; the kernel consumes their fp16 half-plane copies"""
    lib = L.lib()
    fin = fmt_for_dilation(dil)
    S, P = fin
    n = len(dzs)
    arena_dz, offs_dz = _half_arena([to_half_planes(d * scale, S) for d in dzs])
    arena_x, offs_x = _half_arena([to_half_planes(x, S) for x in xs])
    insts = (L.WgradInst * n)()
    for i in range(n):
        insts[i].dz = arena_dz.data_ptr() + 2 * offs_dz[i]
        insts[i].x = arena_x.data_ptr() + 2 * offs_x[i]
    dw = torch.zeros(128, cin_total, ksize, ksize, device="cuda")
    sc = torch.tensor([scale, 1.0 / scale, 0, 0], device="cuda")
    rows = 3 if ksize == 3 else 1
    tasks = (L.WgradTask * rows)()
    for r in range(rows):
        tasks[r] = L.WgradTask(0, n, r, 3 if ksize == 3 else 1, dil, S, P, cin_total, cin0, ksize, dw.data_ptr(),
                               sc.data_ptr())
    L.check(lib.pnmn_debug_launch_wgrad(tasks, rows, insts, n, impl, _stream()))
    torch.cuda.synchronize()
    return dw


def _mkh(shape, gen, scale=1.0):
    return (torch.randn(shape, generator=gen, device="cuda") * scale).half().float()


@pytest.mark.parametrize("impl", [1, 0], ids=["simt", "tcgen05"])
@pytest.mark.parametrize("dil,n_inst", [(1, 3), (2, 1), (4, 2), (8, 2)])
def test_wgrad3x3(impl, dil, n_inst):
    g = torch.Generator(device="cuda").manual_seed(21 + dil)
    dzs = [_mkh((128, 14, 14), g) for _ in range(n_inst)]
    xs = [_mkh((128, 14, 14), g) for _ in range(n_inst)]
    dw = _run_wgrad(impl, dzs, xs, dil, 3, 128, 0, scale=4.0)
    ref = sum(torch.nn.grad.conv2d_weight(x[None], (128, 128, 3, 3), d[None], padding=dil, dilation=dil)
              for x, d in zip(xs, dzs))
    err = _relerr(dw, ref)
    print(f"wgrad impl={impl} dil={dil}: rel err {err:.3e}")
    assert err < 2e-5  # exact fp16 products, fp32 accumulation


@pytest.mark.parametrize("impl", [1, 0], ids=["simt", "tcgen05"])
def test_wgrad_projection_half(impl):
    g = torch.Generator(device="cuda").manual_seed(31)
    dzs = [_mkh((128, 14, 14), g) for _ in range(2)]
    xs = [_mkh((128, 14, 14), g) for _ in range(2)]
    dw = _run_wgrad(impl, dzs, xs, 1, 1, 256, 128)
    ref = sum(torch.einsum("ohw,ihw->oi", d, x) for x, d in zip(xs, dzs))
    err = _relerr(dw[:, 128:, 0, 0], ref)
    print(f"wgrad proj impl={impl}: rel err {err:.3e}")
    assert err < 2e-5
    assert dw[:, :128].abs().max().item() == 0.0


@pytest.mark.parametrize("B,C", [(3, 64), (5, 1024)])
def test_relu_pool_flatten(B, C):
    """ReLU -> MaxPool2d(2,2) -> flatten between the classifier's GEMMs (nmn.py:77-79, nmn_modules.py:250-251):
    the fused pass equals the PyTorch ops bit for bit, forward and backward (ties and all-negative windows included)."""
    from probnmn_clevr_b200.nmn import _ReluPoolFlatten
    g = torch.Generator(device="cuda").manual_seed(5)
    y = torch.randn((B * 196, C), generator=g, device="cuda")
    y[::7] = y[::7].round()          # exact ties inside windows
    y[3::11] = -y[3::11].abs()       # windows without a positive value
    y1 = y.clone().requires_grad_(True)
    out = _ReluPoolFlatten.apply(y1, B)
    y2 = y.clone().requires_grad_(True)
    ref = F.max_pool2d(F.relu(y2).view(B, 14, 14, C).permute(0, 3, 1, 2), 2, 2).contiguous().reshape(B, -1)
    assert torch.equal(out, ref)
    go = torch.randn(out.shape, generator=g, device="cuda")
    out.backward(go)
    ref.backward(go)
    assert torch.equal(y1.grad, y2.grad)


@pytest.mark.parametrize("B,C", [(3, 64), (5, 1024)])
def test_relu_pool_fwd_with_bias(B, C):
    """pnmn_relu_pool_fwd_bias: the conv bias added inside the pooling pass (max(v) + b == max(v + b))"""
    import ctypes
    from probnmn_clevr_b200 import _lib as L
    g = torch.Generator(device="cuda").manual_seed(6)
    y = torch.randn((B * 196, C), generator=g, device="cuda")
    bias = torch.randn(C, generator=g, device="cuda")
    p2 = torch.empty((B, C * 49), device="cuda")
    c2 = torch.empty((B, C * 49), dtype=torch.uint8, device="cuda")
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    L.check(L.lib().pnmn_relu_pool_fwd_bias(ctypes.c_void_p(y.data_ptr()), ctypes.c_void_p(bias.data_ptr()), ctypes.c_void_p(p2.data_ptr()),
                                            ctypes.c_void_p(c2.data_ptr()), B, C, st), "fwd_bias")
    ref = F.max_pool2d(F.relu(y + bias).view(B, 14, 14, C).permute(0, 3, 1, 2), 2, 2).contiguous().reshape(B, -1)
    assert torch.equal(p2, ref)


@pytest.mark.parametrize("M,N,K", [(256, 1024, 50176), (123, 28, 1024), (24108, 1024, 128), (300, 200, 77), (128, 128, 32)])
def test_gemm_split_matches_fp64(M, N, K):
    """pnmn_gemm_split (tcgen05 MMAs over on-the-fly bf16 (hi, lo) splits) against a float64 product, for the three operand
    layouts the classifier uses -- x.w^T (both k-contiguous), g.w (B row-contiguous) and g^T.x (both row-contiguous) --
    with ragged tiles, a split contraction, the bias.  Tolerance: 16 mantissa bits per operand -> ~2e-5 relative L2, far inside
    the 1e-3 bar on the logits; plain fp32 rounding noise of the same product is ~1e-6."""
    from probnmn_clevr_b200.nmn import _gemm
    g = torch.Generator(device="cuda").manual_seed(M + N + K)
    x = torch.randn((M, K), generator=g, device="cuda")
    w = torch.randn((N, K), generator=g, device="cuda") * 0.05
    bias = torch.randn(N, generator=g, device="cuda")
    rel = lambda a, b: float((a.double() - b).norm() / (b.norm() + 1e-300))
    y = _gemm(x, K, 1, w, K, 1, M, N, K, bias=bias)                                  # y = x w^T + b
    assert rel(y, x.double() @ w.double().t() + bias.double()) < 3e-5
    gy = torch.randn((M, N), generator=g, device="cuda") * 1e-4                       # (gradient-sized values)
    dx = _gemm(gy, N, 1, w, 1, K, M, K, N)                                           # dx = g w
    assert rel(dx, gy.double() @ w.double()) < 3e-5
    if M * N * K <= 256 * 1024 * 50176:
        dw = _gemm(gy, 1, N, x, 1, K, N, K, M)                                       # dw = g^T x
        assert rel(dw, gy.double().t() @ x.double()) < 3e-5


def test_classifier_matches_fp32_modules():
    """the classifier on pnmn_gemm_split + the pooling pass against the plain nn.Sequential in IEEE fp32 (cuBLAS / cuDNN, TF32
    off): logits and every gradient"""
    from probnmn_clevr_b200.nmn import NeuralModuleNetwork
    from probnmn_clevr_b200.synthetic import make_nmn_state_dict
    from probnmn_clevr_b200.vocabulary import Vocabulary
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    vocab = Vocabulary.clevr()
    m = NeuralModuleNetwork(vocab)
    m.load_state_dict(make_nmn_state_dict(vocab, 0))
    m = m.cuda()
    g = torch.Generator(device="cuda").manual_seed(3)
    final = torch.randn((37, 128, 14, 14), generator=g, device="cuda").relu_()
    res = {}
    for native in (True, False):
        m.zero_grad(set_to_none=True)
        f = final.clone().requires_grad_(True)
        logits = m._classifier_split(f) if native else m.classifier(f)
        logits.square().sum().backward()
        res[native] = (logits.detach(), f.grad, [p.grad.clone() for p in m.classifier.parameters()])
    rel = lambda a, b: float((a - b).norm() / (b.norm() + 1e-30))
    assert rel(res[True][0], res[False][0]) < 1e-4
    # gradients: a 2x2 window whose two largest pre-activations differ by less than the products' rounding noise routes its
    # gradient to the other pixel (max-pool is discontinuous there) -- a handful of such windows out of 37 * 50176 give
    # ~2e-3 relative L2 on d(input) and on the conv's gradients; everything downstream of the pooling (both Linear layers)
    # agrees to 1e-4
    assert rel(res[True][1], res[False][1]) < 1e-2
    grads = list(zip(res[True][2], res[False][2]))
    for a, b in grads[:2]:
        assert rel(a, b) < 1e-2
    for a, b in grads[2:]:
        assert rel(a, b) < 1e-4


def test_nchw_to_planes_writes_the_fp16_shadow():
    """Stem input staging: NCHW fp32 features -> fp16 half planes [C/8][256 slots][8] behind the (unwritten) fp32 planes of
    the unit; padding slots stay zero (extract_features.py:102-131 gives NCHW fp32; executor.h describes the planes)."""
    import ctypes
    from probnmn_clevr_b200 import _lib as L
    B, C = 3, 256
    g = torch.Generator(device="cuda").manual_seed(13)
    x = torch.randn((B, C, 14, 14), generator=g, device="cuda").relu_()
    unit_floats = (C // 4) * 256 * 4 * 3 // 2
    dst = torch.zeros(B * unit_floats, device="cuda")
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    L.check(L.lib().pnmn_debug_nchw_to_planes(ctypes.c_void_p(x.data_ptr()), ctypes.c_void_p(dst.data_ptr()), B, C,
                                              unit_floats, st), "nchw_to_planes")
    units = dst.view(B, unit_floats)
    assert units[:, : (C // 4) * 256 * 4].abs().max().item() == 0.0          # fp32 planes: untouched
    shadow = units[:, (C // 4) * 256 * 4:].contiguous().view(torch.float16).view(B, C // 8, 16, 16, 8)
    # tf32 rounding (ties away from zero on the 13 dropped bits), then fp16
    xi = x.view(torch.int32)
    tf32 = ((xi + 0x1000) & ~0x1FFF).view(torch.float32)
    want = tf32.half().view(B, C // 8, 8, 14, 14).permute(0, 1, 3, 4, 2)
    assert torch.equal(shadow[:, :, :14, :14, :], want)
    assert shadow[:, :, 14:].abs().max().item() == 0.0 and shadow[:, :, :, 14:].abs().max().item() == 0.0


@pytest.mark.parametrize("with_answers", [True, False], ids=["cross_entropy", "max_logprob"])
def test_answer_loss_head_matches_torch(with_answers):
    """pnmn_answer_loss_forward / _backward against the eager ops of the reference's answer head (nmn.py:245-269): first-maximum
    predictions (ties included), @@UNKNOWN@@ and the constant 3.33 for invalid rows, per-row losses, correct count, and the
    gradient of a weighted sum of the losses (zero rows for invalid programs)"""
    from probnmn_clevr_b200.nmn import _AnswerLoss
    torch.manual_seed(3)
    B, A, unknown = 77, 28, 1
    logits = torch.randn(B, A, device="cuda") * 3
    logits[5, 7] = logits[5, 3] = logits[5].max() + 1.0          # a tie: the lower index wins
    answers = torch.randint(0, A, (B,), device="cuda")
    answers[5] = 3
    xin = torch.arange(B, dtype=torch.int64, device="cuda") * 4096
    bad = torch.zeros(B, dtype=torch.bool, device="cuda")
    bad[[2, 5 + 1, 40, 76]] = True
    xin[bad] = -1
    blob = torch.cat([torch.zeros(256, dtype=torch.uint8, device="cuda"), xin.view(torch.uint8)])
    w = torch.rand(B, device="cuda")
    x = logits.clone().requires_grad_(True)
    correct = torch.zeros((), dtype=torch.int64, device="cuda") if with_answers else None
    pred, loss = _AnswerLoss.apply(x, answers if with_answers else None, [(blob, 256, 0, B)], unknown, correct)
    (loss * w).sum().backward()

    y = logits.clone().requires_grad_(True)
    logp = torch.log_softmax(y, dim=-1)
    best, ref_pred = torch.max(logp, dim=1)
    ref_pred = ref_pred.masked_fill(bad, unknown)
    ref_loss = torch.nn.functional.cross_entropy(y, answers, reduction="none") if with_answers else -best
    ref_loss = ref_loss.masked_fill(bad, 3.33)
    (ref_loss * w).sum().backward()
    assert torch.equal(pred, ref_pred) and int(pred[5]) == 3
    assert torch.allclose(loss, ref_loss, rtol=1e-6, atol=1e-6)
    assert torch.allclose(x.grad, y.grad, rtol=1e-5, atol=1e-7)
    assert float(x.grad[bad].abs().max()) == 0.0
    if with_answers:
        assert int(correct) == int((ref_pred == answers).sum())
