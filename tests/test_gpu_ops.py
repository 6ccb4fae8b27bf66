"""-m gpu parity tests of the operator-level C-ABI entry points (K1, K4, K5) against the oracle
(log-mel) and a plain PyTorch fp32 reference (LayerNorm, GEMM)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _bf(x):
    return x.to(torch.bfloat16)


@pytest.fixture(scope="module")
def ops(cuda_device):
    from turbo_whisper_workspace_b200 import ops
    return ops


# ------------------------------------------------------------------ K1 log-mel
@pytest.mark.parametrize("n_valid", [480000, 160000, 12345, 1, 0])
def test_logmel_vs_oracle(ops, cuda_device, n_valid):
    from oracle import logmel_ref as L
    rng = np.random.default_rng(0)
    B = 3
    pcm = (0.1 * rng.standard_normal((B, 480000))).astype(np.float32)
    # clip 1: amplitude-modulated noise + a sine sweep so the max-8 clamp is exercised
    t = np.arange(480000) / 16000.0
    pcm[1] = (0.3 * np.sin(2 * np.pi * (200 + 100 * t) * t) * (t < 20) + 1e-4 * pcm[1]).astype(np.float32)
    nv = np.array([n_valid, 480000, max(n_valid // 2, 0)], dtype=np.int32)
    lm = ops.LogMel(cuda_device, max_batch=4)
    d_pcm = torch.from_numpy(pcm).to(cuda_device)
    d_nv = torch.from_numpy(nv).to(cuda_device)
    out = torch.empty(B, 128, 3000, dtype=torch.float32, device=cuda_device)
    out_t = torch.zeros(B, 3002, 128, dtype=torch.bfloat16, device=cuda_device)
    lm(d_pcm, d_nv, out_f32=out, out_t=out_t, out_t_row_off=1)
    torch.cuda.synchronize()
    got = out.cpu().numpy()
    for b in range(B):
        ref = L.log_mel(pcm[b, :nv[b]])
        # tolerance: 1e-4 relative to the feature scale (north star: log-mel within 1e-4 relative)
        np.testing.assert_allclose(got[b], ref, rtol=1e-4, atol=1e-4)
    # the bf16 time-major copy is the same data, transposed, rounded to bf16, at row offset 1
    tr = out_t[:, 1:3001, :].float().transpose(1, 2)
    assert torch.equal(tr, out.to(torch.bfloat16).float())
    assert float(out_t[:, 0].abs().max()) == 0.0 and float(out_t[:, 3001].abs().max()) == 0.0


def test_logmel_golden_kat(ops, cuda_device):
    """Known-answer values measured from transformers 5.5.0."""
    pcm = (0.1 * np.random.default_rng(0).standard_normal(480000)).astype(np.float32)
    lm = ops.LogMel(cuda_device, max_batch=1)
    out = torch.empty(1, 128, 3000, dtype=torch.float32, device=cuda_device)
    lm(torch.from_numpy(pcm)[None].to(cuda_device), None, out_f32=out)
    x = out[0].cpu().numpy()
    assert abs(x.mean() - 0.597388) < 1e-5
    np.testing.assert_allclose(x[0, :4], [0.45651639, 0.56622541, 0.60694182, 0.39072639], atol=2e-5)
    np.testing.assert_allclose(x[127, -4:], [0.74227071, 0.59086835, 0.66143847, 0.55686784], atol=2e-5)
    assert abs(x.max() - 0.942317) < 2e-5 and abs(x.min() + 0.501018) < 2e-5


# ------------------------------------------------------------------ K4 LayerNorm
@pytest.mark.parametrize("rows,cols", [(1, 1280), (37, 1280), (1500 * 3, 1280), (5, 256), (9, 2048)])
def test_layernorm(ops, cuda_device, rows, cols):
    g = torch.Generator(device="cpu").manual_seed(1)
    x = (torch.randn(rows, cols, generator=g) * 3 + 0.5).to(cuda_device)
    w = torch.randn(cols, generator=g).to(cuda_device)
    b = torch.randn(cols, generator=g).to(cuda_device)
    got = ops.layernorm(x, w, b).float()
    ref = torch.nn.functional.layer_norm(x, (cols,), w, b, 1e-5)
    torch.testing.assert_close(got, ref, rtol=1e-2, atol=1e-2)  # bf16 output rounding (2^-8 relative)


# ------------------------------------------------------------------ K5 GEMM
def _gemm_ref(a, w, bias=None, act=0, resid=None):
    y = a.float() @ w.float().t()
    if bias is not None:
        y = y + bias
    if act == 1:
        y = torch.nn.functional.gelu(y)
    if resid is not None:
        y = y + resid
    return y


@pytest.mark.parametrize("M,N,K", [(128, 256, 64), (128, 256, 256), (300, 512, 1280), (1500, 1280, 1280),
                                    (1000, 3840, 1280), (777, 1280, 5120), (129, 264, 72), (4500, 5120, 1280)])
def test_gemm_plain(ops, cuda_device, M, N, K):
    g = torch.Generator(device="cpu").manual_seed(M + N + K)
    a = _bf(torch.randn(M, K, generator=g)).to(cuda_device)
    w = _bf(torch.randn(N, K, generator=g) / K ** 0.5).to(cuda_device)
    out = torch.full((M, N), float("nan"), dtype=torch.bfloat16, device=cuda_device)
    ops.gemm(a, w, rows=M, out=out)
    torch.cuda.synchronize()
    ref = _gemm_ref(a, w)
    # fp32 accumulation of exact bf16 products; only the bf16 output rounding (2^-9) remains
    torch.testing.assert_close(out.float(), ref, rtol=8e-3, atol=8e-3)


@pytest.mark.parametrize("act,use_resid,out_f32", [(0, False, True), (1, False, False), (0, True, True), (1, True, True)])
def test_gemm_epilogues(ops, cuda_device, act, use_resid, out_f32):
    M, N, K = 700, 1280, 640
    g = torch.Generator(device="cpu").manual_seed(7)
    a = _bf(torch.randn(M, K, generator=g)).to(cuda_device)
    w = _bf(torch.randn(N, K, generator=g) / K ** 0.5).to(cuda_device)
    bias = torch.randn(N, generator=g).to(cuda_device)
    resid = torch.randn(M, N, generator=g).to(cuda_device) if use_resid else None
    out = torch.zeros(M, N, dtype=torch.float32 if out_f32 else torch.bfloat16, device=cuda_device)
    if use_resid and out_f32:  # in-place residual stream update, as the encoder uses it
        out.copy_(resid)
        ops.gemm(a, w, rows=M, bias=bias, act=act, resid=out, resid_ld=N, out=out)
    else:
        ops.gemm(a, w, rows=M, bias=bias, act=act, resid=resid, resid_ld=N, out=out)
    ref = _gemm_ref(a, w, bias, act, resid)
    tol = 2e-4 if out_f32 else 8e-3
    torch.testing.assert_close(out.float(), ref, rtol=tol, atol=tol)


def test_gemm_conv_views(ops, cuda_device):
    """Both Conv1d layers of the stem as implicit GEMMs over time-major activations."""
    g = torch.Generator(device="cpu").manual_seed(11)
    B, T, Cin, C = 2, 3000, 128, 256
    x = _bf(torch.randn(B, Cin, T, generator=g))                     # [B, C, T] like HF input_features
    w1 = _bf(torch.randn(C, Cin, 3, generator=g) / (3 * Cin) ** 0.5)
    b1 = torch.randn(C, generator=g)
    w2 = _bf(torch.randn(C, C, 3, generator=g) / (3 * C) ** 0.5)
    b2 = torch.randn(C, generator=g)
    pos = torch.randn(T // 2, C, generator=g)
    F = torch.nn.functional
    h1_ref = F.gelu(F.conv1d(x.float(), w1.float(), b1, padding=1))                     # [B, C, T]
    h1_bf = _bf(h1_ref)
    y_ref = F.gelu(F.conv1d(h1_bf.float(), w2.float(), b2, stride=2, padding=1)).transpose(1, 2) + pos  # [B,T/2,C]

    dev = cuda_device
    rows_t = 1 + T + T + 1
    xt = torch.zeros(B, rows_t, Cin, dtype=torch.bfloat16, device=dev)
    xt[:, 1:T + 1] = x.transpose(1, 2).to(dev)
    w1r = w1.permute(0, 2, 1).reshape(C, 3 * Cin).contiguous().to(dev)   # [o, k*Cin + c]
    w2r = w2.permute(0, 2, 1).reshape(C, 3 * C).contiguous().to(dev)
    h1 = torch.zeros(B, T + 2, C, dtype=torch.bfloat16, device=dev)
    ops.gemm(xt, w1r, rows=T, batches=B, a_row_stride=Cin, a_batch_stride=rows_t * Cin, a_rows=2 * T,
             bias=b1.to(dev), act=1, out=h1, out_ld=C, out_batch_rows=T + 2, out_row_off=1)
    torch.testing.assert_close(h1[:, 1:T + 1].float().cpu(), h1_ref.transpose(1, 2), rtol=1e-2, atol=1e-2)
    assert float(h1[:, 0].abs().max()) == 0 and float(h1[:, T + 1].abs().max()) == 0
    # feed the reference's bf16 h1 into conv2 so the check isolates the second GEMM
    h1.zero_()
    h1[:, 1:T + 1] = h1_bf.transpose(1, 2).to(dev)
    y = torch.zeros(B * (T // 2), C, dtype=torch.float32, device=dev)
    ops.gemm(h1, w2r, rows=T // 2, batches=B, a_row_stride=2 * C, a_batch_stride=(T + 2) * C, a_rows=T // 2,
             bias=b2.to(dev), act=1, resid=pos.to(dev), resid_ld=C, resid_batch_rows=0, out=y, out_ld=C,
             out_batch_rows=T // 2)
    torch.testing.assert_close(y.view(B, T // 2, C).cpu(), y_ref, rtol=2e-3, atol=2e-3)
    # a_row_off: per-batch row offset into the A view (rows >= 1 see the true left neighbour)
    seek = torch.tensor([0, 1234], dtype=torch.int32, device=dev)
    h1s = torch.zeros(B, T + 2, C, dtype=torch.bfloat16, device=dev)
    ops.gemm(xt, w1r, rows=T, batches=B, a_row_stride=Cin, a_batch_stride=rows_t * Cin, a_rows=2 * T,
             a_row_off=seek, bias=b1.to(dev), act=1, out=h1s, out_ld=C, out_batch_rows=T + 2, out_row_off=1)
    xs = torch.zeros_like(x)
    xs[1, :, :T - 1234] = x[1, :, 1234:]
    xs[0] = x[0]
    ref_s = F.gelu(F.conv1d(xs.float(), w1.float(), b1, padding=1)).transpose(1, 2)
    torch.testing.assert_close(h1s[:, 2:T + 1].float().cpu(), ref_s[:, 1:], rtol=1e-2, atol=1e-2)


# ------------------------------------------------------------------ K6 encoder attention
@pytest.mark.parametrize("B,T,H", [(1, 128, 1), (1, 256, 2), (2, 1500, 4), (1, 100, 3), (3, 1500, 20)])
def test_attention_enc(ops, cuda_device, B, T, H):
    g = torch.Generator(device="cpu").manual_seed(B * 1000 + T + H)
    D = H * 64
    qkv = torch.randn(B * T, 3 * D, generator=g)
    qkv[:, :D] *= 0.35   # q already carries the 1/8 scale in the engine; keep softmax peaky but finite
    qkv = _bf(qkv).to(cuda_device)
    out = ops.attention_enc(qkv, B, T, H)
    torch.cuda.synchronize()
    q, k, v = [t.float().view(B, T, H, 64).transpose(1, 2) for t in qkv.split(D, dim=1)]
    ref = (torch.softmax(q @ k.transpose(-1, -2), dim=-1) @ v).transpose(1, 2).reshape(B * T, D)
    # P is rounded to bf16 before PV (2^-9 relative per term) and the output is stored in bf16
    torch.testing.assert_close(out.float(), ref, rtol=2e-2, atol=2e-2)
    assert float((out.float() - ref).abs().mean()) < 2e-3


def test_shift_frames(ops, cuda_device):
    g = torch.Generator(device="cpu").manual_seed(5)
    src = torch.zeros(3, 3002, 128, dtype=torch.bfloat16)
    src[:, 1:3001] = _bf(torch.randn(3, 3000, 128, generator=g))
    src = src.to(cuda_device)
    dst = torch.full_like(src[:2], 7.0)
    dst[:, 0] = 0
    dst[:, 3001] = 0
    seek = torch.tensor([1234, 0], dtype=torch.int32, device=cuda_device)
    rows = torch.tensor([2, 0], dtype=torch.int32, device=cuda_device)
    ops.shift_frames(src, dst, seek, src_row=rows)
    assert torch.equal(dst[0, 1:3001 - 1234], src[2, 1235:3001])
    assert float(dst[0, 3001 - 1234:].abs().max()) == 0.0
    assert torch.equal(dst[1], src[0])


# ------------------------------------------------------------------ K0 audio ingest (convert + down-mix + resample)
@pytest.mark.parametrize("orig,ch,dtype", [(48000, 1, "f32"), (44100, 2, "i16"), (8000, 1, "i16"), (22050, 3, "f32")])
def test_resample_matches_torchaudio(ops, cuda_device, orig, ch, dtype):
    """tw_resample vs torchaudio.functional.resample on the channel mean (the reference's preprocess branch,
    $TF/pipelines/automatic_speech_recognition.py:394-407).  fp32 dot products of ~35 taps in a different summation order than conv1d: 1e-5 absolute on
    |x| <= 0.8 (a third of a 16-bit LSB)."""
    import torchaudio.functional as AF
    g = torch.Generator().manual_seed(orig + ch)
    n = orig * 3 + 1234
    x = torch.rand(n, ch, generator=g) * 1.6 - 0.8
    if dtype == "i16":
        xi = (x * 32767).round().to(torch.int16)
        host = xi.float() / 32768.0
        dev_in = xi.to(cuda_device)
    else:
        host = x
        dev_in = x.to(cuda_device)
    want = AF.resample(host.mean(dim=1), orig, 16000)
    rs = ops.Resampler(orig, 16000, cuda_device)
    got = rs(dev_in if ch > 1 else dev_in[:, 0].contiguous()).cpu()
    assert got.shape == want.shape
    assert float((got - want).abs().max()) < 1e-5
    # empty and tiny inputs
    assert rs(torch.zeros(0, dtype=torch.float32, device=cuda_device)).numel() == 0
    tiny = torch.ones(5, dtype=torch.float32, device=cuda_device)
    assert torch.allclose(rs(tiny).cpu(), AF.resample(torch.ones(5), orig, 16000), atol=1e-5)
