"""Parity of each sm_100a kernel (through the C ABI) against the oracle / a plain fp32 torch restatement.
Integer/fp32-exact work (sampler update, tile blend) must be bit-identical; bf16 tensor-core work is checked
against fp32 math on the same bf16-rounded inputs with a stated tolerance."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

BF16_TOL = 1e-2  # max-abs error relative to max-abs of the reference, bf16 outputs (8 mantissa bits)


def rel(a, b):
    a, b = a.float(), b.float()
    return ((a - b).abs().max() / (b.abs().max() + 1e-12)).item()


@pytest.fixture(scope="module")
def ops(cuda_lib):
    from tair_b200 import ops as o
    return o


def rn(*shape, scale=1.0, seed=0):
    g = torch.Generator(device="cuda").manual_seed(seed)
    return torch.randn(*shape, device="cuda", generator=g) * scale


@pytest.mark.parametrize("M,N,K", [(128, 128, 64), (4096, 320, 320), (1000, 320, 320), (1232, 640, 1024),
                                   (16, 1280, 320), (300, 97, 256), (77, 2, 256), (5000, 1920, 640), (65, 8, 72)])
def test_gemm_shapes(ops, M, N, K):
    a, w = rn(M, K).bfloat16(), rn(N, K, scale=K ** -0.5, seed=1).bfloat16()
    out = ops.gemm(a, w)
    assert rel(out, a.float() @ w.float().t()) < BF16_TOL


def test_gemm_epilogues(ops):
    M, N, K = 2048, 640, 320
    a, w = rn(M, K).bfloat16(), rn(N, K, scale=K ** -0.5, seed=1).bfloat16()
    bias, res, rg = rn(N, seed=2), rn(M, N, seed=3).bfloat16(), rn(M // 512, N, seed=4)
    base = a.float() @ w.float().t() + bias
    assert rel(ops.gemm(a, w, bias=bias, residual=res), base + res.float()) < BF16_TOL
    assert rel(ops.gemm(a, w, bias=bias, rowgroup=rg, rows_per_group=512), base + rg.repeat_interleave(512, 0)) < BF16_TOL
    assert rel(ops.gemm(a, w, bias=bias, act=ops.ACT_GELU), F.gelu(base)) < BF16_TOL
    assert rel(ops.gemm(a, w, bias=bias, act=ops.ACT_SILU), F.silu(base)) < BF16_TOL
    assert rel(ops.gemm(a, w, bias=bias, act=ops.ACT_RELU), F.relu(base)) < BF16_TOL
    assert rel(ops.gemm(a, w, bias=bias, out_dtype=torch.float32), base) < 1e-5
    # strided output / strided A (column slices of wider buffers)
    wide = torch.zeros(M, N + 64, device="cuda", dtype=torch.bfloat16)
    ops.gemm(a, w, bias=bias, out=wide[:, 64:])
    assert rel(wide[:, 64:], base) < BF16_TOL and wide[:, :64].abs().max() == 0
    a_wide = rn(M, K + 128, seed=7).bfloat16()
    assert rel(ops.gemm(a_wide[:, 128:], w), a_wide[:, 128:].float() @ w.float().t()) < BF16_TOL


@pytest.mark.parametrize("M,N,K", [(4096, 256, 256), (5000, 384, 256), (3000, 1536, 256), (2500, 1024, 256),
                                   (4096, 320, 320), (2048, 640, 640), (1000, 1280, 1280), (300, 96, 64),
                                   (9472, 160, 256), (1111, 224, 128)])
def test_gemm_lean_epilogue_variants(ops, M, N, K):
    """The specialised TMA-store epilogues (csrc/gemm_tc.cu: epilogue_tile_tma_lean): plain, bias, bias + activation,
    bias + residual, residual alone, and bf16 row-group rows in both addressing forms, at the N tiles the tuner may pick
    (N % 32 == 0 selects them; ragged M exercises the row clipping)."""
    a, w = rn(M, K).bfloat16(), rn(N, K, scale=K ** -0.5, seed=1).bfloat16()
    bias, res = rn(N, seed=2), rn(M, N, seed=3).bfloat16()
    raw = a.float() @ w.float().t()
    base = raw + bias
    assert rel(ops.gemm(a, w), raw) < BF16_TOL
    assert rel(ops.gemm(a, w, bias=bias), base) < BF16_TOL
    assert rel(ops.gemm(a, w, bias=bias, act=ops.ACT_RELU), F.relu(base)) < BF16_TOL
    assert rel(ops.gemm(a, w, bias=bias, act=ops.ACT_GELU), F.gelu(base)) < BF16_TOL
    assert rel(ops.gemm(a, w, bias=bias, act=ops.ACT_SILU), F.silu(base)) < BF16_TOL
    assert rel(ops.gemm(a, w, bias=bias, residual=res), base + res.float()) < BF16_TOL
    assert rel(ops.gemm(a, w, residual=res), raw + res.float()) < BF16_TOL
    # residual aliasing the output (x = x + f(x), the transformer blocks' pattern)
    buf = res.clone()
    ops.gemm(a, w, bias=bias, residual=buf, out=buf)
    assert rel(buf, base + res.float()) < BF16_TOL
    # bf16 row groups: one row per group of consecutive output rows / periodic rows
    per = 25
    G = (M + per - 1) // per
    rg = rn(G, N, seed=4).bfloat16()
    want = raw + rg.float().repeat_interleave(per, 0)[:M]
    assert rel(ops.gemm(a, w, rowgroup=rg, rows_per_group=per), want) < BF16_TOL
    P = 100
    rgp = rn(P, N, seed=5).bfloat16()
    want = raw + rgp.float()[torch.arange(M, device="cuda") % P] + bias
    assert rel(ops.gemm(a, w, bias=bias, rowgroup=rgp, rows_per_group=-P), want) < BF16_TOL
    # the fp32 form of the same rows goes through the generic epilogue and must agree to bf16 rounding of the rows
    got32 = ops.gemm(a, w, bias=bias, rowgroup=rgp.float(), rows_per_group=-P)
    assert rel(got32, want) < BF16_TOL


def test_gemm_row_add_full_size_properties(ops):
    """The K = 256 GEMMs of a B=16 TESTR encoder layer (151552 rows): the row-add epilogue is linear in what it adds -
    gemm(a, w, residual=r) - gemm(a, w) reproduces r, the periodic bf16 row group reproduces its table on every image -
    and tiles past the last full one (151552 = 1184 x 128 exactly; 151500 is ragged) are clipped, not overrun."""
    K = N = 256
    for M in (151552, 151500):
        a, w = rn(M, K, seed=7).bfloat16(), rn(N, K, scale=K ** -0.5, seed=8).bfloat16()
        plain = ops.gemm(a, w).float()
        r = rn(M, N, seed=9).bfloat16()
        canary = torch.full((M + 128, N), 3.0, device="cuda", dtype=torch.bfloat16)
        ops.gemm(a, w, residual=r, out=canary[:M])
        assert (canary[M:] == 3.0).all()
        d = canary[:M].float() - plain
        assert (d - r.float()).abs().max() <= 2 ** -7 * (plain.abs().max() + r.float().abs().max())   # two bf16 roundings
        S = 9472
        table = rn(S, N, seed=10).bfloat16()
        got = ops.gemm(a, w, rowgroup=table, rows_per_group=-S).float() - plain
        want = table.float()[torch.arange(M, device="cuda") % S]
        assert (got - want).abs().max() <= 2 ** -7 * (plain.abs().max() + want.abs().max())


def test_gemm_bf16_rowgroup_limits(ops):
    """bf16 row groups exist only in the row-add epilogue: combinations it does not cover fail loudly."""
    from tair_b200._lib import TairError
    M, N, K = 512, 256, 256
    a, w = rn(M, K).bfloat16(), rn(N, K, scale=K ** -0.5, seed=1).bfloat16()
    rg = rn(4, N, seed=2).bfloat16()
    for kw in (dict(act=ops.ACT_RELU), dict(residual=rn(M, N, seed=3).bfloat16()), dict(out_dtype=torch.float32)):
        with pytest.raises(TairError):
            ops.gemm(a, w, rowgroup=rg, rows_per_group=128, **kw)
    w2 = rn(72, K, scale=K ** -0.5, seed=1).bfloat16()       # N % 32 != 0
    with pytest.raises(TairError):
        ops.gemm(a, w2, rowgroup=rn(4, 72, seed=2).bfloat16(), rows_per_group=128)


def test_gemm_geglu_matches_chunked_reference(ops):
    from tair_b200.model.attention import interleave_geglu
    M, C = 1000, 320
    a = rn(M, C).bfloat16()
    w, b = rn(8 * C, C, scale=C ** -0.5, seed=1), rn(8 * C, seed=2)
    wi, bi = interleave_geglu(w.bfloat16(), b)
    out = ops.gemm(a, wi, bias=bi, act=ops.ACT_GEGLU)
    val, gate = (a.float() @ w.bfloat16().float().t() + b).chunk(2, dim=-1)
    assert rel(out, val * F.gelu(gate)) < BF16_TOL


@pytest.mark.parametrize("B,H,Cin,Cout,stride", [(1, 16, 64, 64, 1), (2, 64, 320, 320, 1), (2, 32, 640, 640, 1),
                                                  (3, 8, 1280, 1280, 1), (1, 8, 2560, 1280, 1), (2, 64, 320, 4, 1),
                                                  (2, 64, 320, 320, 2), (2, 16, 1280, 1280, 2), (1, 4, 64, 64, 1)])
def test_conv3x3(ops, B, H, Cin, Cout, stride):
    x = rn(B, Cin, H, H).bfloat16()
    w = rn(Cout, Cin, 3, 3, scale=(9 * Cin) ** -0.5, seed=1).bfloat16()
    bias = rn(Cout, seed=2)
    out = ops.conv3x3(x.permute(0, 2, 3, 1).contiguous(), w.permute(0, 2, 3, 1).reshape(Cout, -1).contiguous(),
                      stride=stride, bias=bias)
    ref = F.conv2d(x.float(), w.float(), bias, stride=stride, padding=1).permute(0, 2, 3, 1)
    assert rel(out, ref) < BF16_TOL


def test_conv3x3_fused_epilogue(ops):
    B, H, C = 2, 32, 320
    x = rn(B, C, H, H).bfloat16()
    w = rn(C, C, 3, 3, scale=(9 * C) ** -0.5, seed=1).bfloat16()
    bias, emb, res = rn(C, seed=2), rn(B, C, seed=3), rn(B, H, H, C, seed=4).bfloat16()
    out = ops.conv3x3(x.permute(0, 2, 3, 1).contiguous(), w.permute(0, 2, 3, 1).reshape(C, -1).contiguous(), bias=bias,
                      rowgroup=emb, rows_per_group=H * H, residual=res)
    ref = F.conv2d(x.float(), w.float(), bias, padding=1) + emb[:, :, None, None]
    assert rel(out, ref.permute(0, 2, 3, 1) + res.float()) < BF16_TOL


@pytest.mark.parametrize("B,H,Lq,Lk", [(1, 1, 128, 128), (2, 5, 1024, 1024), (2, 10, 1024, 77), (3, 20, 64, 64),
                                       (3, 20, 64, 77), (1, 3, 200, 333), (1, 5, 4096, 4096),
                                       # short contexts take the key/value-stationary kernel: many / few / ragged query tiles
                                       (4, 5, 4096, 77), (2, 20, 256, 77), (1, 5, 200, 50), (5, 2, 1300, 80), (2, 3, 128, 1)])
def test_attention(ops, B, H, Lq, Lk):
    C = H * 64
    q, k, v = rn(B * Lq, C, scale=1.5).bfloat16(), rn(B * Lk, C, scale=1.5, seed=1).bfloat16(), rn(B * Lk, C, seed=2).bfloat16()
    out = ops.attention(q, k, v, B=B, H=H, Lq=Lq, Lk=Lk)
    def heads(t, L):
        return t.float().view(B, L, H, 64).transpose(1, 2)
    ref = F.scaled_dot_product_attention(heads(q, Lq), heads(k, Lk), heads(v, Lk)).transpose(1, 2).reshape(B * Lq, C)
    assert torch.isfinite(out.float()).all() and rel(out, ref) < BF16_TOL


def test_attention_fused_qkv_slices(ops):
    B, H, L = 2, 5, 256
    C = H * 64
    qkv = rn(B * L, 3 * C).bfloat16()
    out = ops.attention(qkv[:, :C], qkv[:, C:2 * C], qkv[:, 2 * C:], B=B, H=H, Lq=L, Lk=L)
    def heads(t):
        return t.float().reshape(B, L, H, 64).transpose(1, 2)
    ref = F.scaled_dot_product_attention(heads(qkv[:, :C]), heads(qkv[:, C:2 * C]), heads(qkv[:, 2 * C:]))
    assert rel(out, ref.transpose(1, 2).reshape(B * L, C)) < BF16_TOL


@pytest.mark.parametrize("B,HW,C,eps,act", [(2, 4096, 320, 1e-5, "silu"), (2, 1024, 640, 1e-6, "none"),
                                            (3, 64, 2560, 1e-5, "silu"), (2, 1024, 1920, 1e-5, "silu"),
                                            (2, 4096, 256, 1e-5, "gelu"), (2, 1024, 64, 1e-5, "silu"),
                                            (1, 256, 192, 1e-5, "silu"), (2, 4096, 960, 1e-5, "silu")])
def test_groupnorm(ops, B, HW, C, eps, act):
    x = (rn(B, HW, C) * 1.7 + 0.4).bfloat16()
    g, b = 1 + 0.1 * rn(C, seed=1), 0.1 * rn(C, seed=2)
    a = {"silu": ops.ACT_SILU, "none": ops.ACT_NONE, "gelu": ops.ACT_GELU}[act]
    out = ops.groupnorm(x, g, b, eps=eps, act=a)
    ref = F.group_norm(x.float().transpose(1, 2), 32, g, b, eps).transpose(1, 2)
    ref = {"silu": F.silu, "none": lambda t: t, "gelu": F.gelu}[act](ref)
    assert rel(out, ref) < BF16_TOL


@pytest.mark.parametrize("M,C", [(4096, 320), (1000, 640), (333, 1280), (9472, 256), (7, 2048), (1001, 256), (3, 64),
                                 (2501, 512), (77, 384), (5, 8)])
def test_layernorm(ops, M, C):
    x = (rn(M, C) * 2 + 0.3).bfloat16()
    g, b = 1 + 0.1 * rn(C, seed=1), 0.1 * rn(C, seed=2)
    assert rel(ops.layernorm(x, g, b), F.layer_norm(x.float(), (C,), g, b, 1e-5)) < BF16_TOL


def test_sampler_update_bit_exact_vs_oracle(ops):
    from oracle import sampler
    tabs = sampler.tables_to_torch(sampler.make_schedule(sampler.diffusion_betas(), 50), "cuda")
    order = ["sqrt_alphas_cumprod", "sqrt_one_minus_alphas_cumprod", "posterior_mean_coef1", "posterior_mean_coef2",
             "posterior_variance"]
    B = 16
    x, v, vu, nz = (rn(B, 4, 64, 64, seed=s) for s in range(4))
    for idx in (49, 25, 1, 0):
        t = torch.full((B,), idx, device="cuda", dtype=torch.long)
        ref, ref0 = sampler.p_sample_update(tabs, x, v, t, nz)
        x0 = torch.empty_like(x)
        out = ops.sampler_update(x, v, nz, t, [tabs[k] for k in order], pred_x0=x0)
        assert torch.equal(out, ref) and torch.equal(x0, ref0), idx
        ref_cfg, _ = sampler.p_sample_update(tabs, x, v, t, nz, v_uncond=vu, cfg_scale=4.0)
        assert torch.equal(ops.sampler_update(x, v, nz, t, [tabs[k] for k in order], v_uncond=vu, cfg_scale=4.0), ref_cfg)
    t = torch.randint(0, 50, (B,), device="cuda")  # per-sample indices (extract_into_tensor semantics)
    assert torch.equal(ops.sampler_update(x, v, nz, t, [tabs[k] for k in order]), sampler.p_sample_update(tabs, x, v, t, nz)[0])


def test_layout_and_elementwise(ops):
    x = rn(3, 4, 64, 64)
    nhwc = ops.nchw_to_nhwc(x, 64)
    assert nhwc.shape == (3, 64, 64, 64) and nhwc[..., 4:].abs().max() == 0
    assert torch.equal(nhwc[..., :4], x.permute(0, 2, 3, 1).bfloat16())
    f = rn(2, 37, 24, 320).bfloat16()
    assert torch.equal(ops.nhwc_to_nchw(f), f.float().permute(0, 3, 1, 2))
    assert torch.equal(ops.nhwc_to_nchw(nhwc, 4), x.bfloat16().float())
    a, b, c = rn(2, 8, 8, 640).bfloat16(), rn(2, 8, 8, 320, seed=1).bfloat16(), rn(2, 8, 8, 320, seed=2).bfloat16()
    assert torch.equal(ops.concat_add(a, b, c), torch.cat([a, (b.float() + c.float()).bfloat16()], -1))
    assert torch.equal(ops.concat_add(a, b), torch.cat([a, b], -1))
    assert torch.equal(ops.add(b, c), (b.float() + c.float()).bfloat16())
    up = ops.upsample2x(b)
    assert torch.equal(up, F.interpolate(b.permute(0, 3, 1, 2).float(), scale_factor=2, mode="nearest").permute(0, 2, 3, 1).bfloat16())
    t = torch.tensor([0, 20, 500, 999], device="cuda")
    from oracle.unet import timestep_embedding
    assert (ops.timestep_embedding(t, 320).float() - timestep_embedding(t, 320)).abs().max() < 1e-2


def test_msda_forward_vs_oracle_and_fixture(ops, golden):
    from oracle import msda, weights
    g = golden("msda_case.npz")
    shapes = [tuple(int(v) for v in r) for r in g["shapes"]]
    S = sum(h * w for h, w in shapes)
    B, M, D, Lq, L, P = 2, 8, 32, 24, 3, 4
    value = weights.seeded_randn((B, S, M, D), 21).cuda()
    loc = (torch.rand((B, Lq, M, L, P, 2), generator=torch.Generator().manual_seed(22)) * 1.4 - 0.2).cuda()
    w = torch.softmax(weights.seeded_randn((B, Lq, M, L * P), 23), -1).view(B, Lq, M, L, P).cuda()
    shp = torch.tensor(shapes, device="cuda", dtype=torch.long)
    start = torch.cat([shp.new_zeros(1), (shp[:, 0] * shp[:, 1]).cumsum(0)[:-1]])
    out = ops.msda_forward(value, shp, start, loc, w)
    assert (out.cpu().numpy() - g["out"]).__abs__().max() < 1e-5           # reference fixture, fp32 mode
    # TESTR geometry (4 levels, 9472 tokens) against the oracle, fp32 and bf16 value
    shapes = [(16, 16), (32, 32), (64, 64), (64, 64)]
    S = 9472
    B, Lq = 2, 700
    value, loc = rn(B, S, 8, 32), torch.rand(B, Lq, 8, 4, 4, 2, device="cuda") * 1.2 - 0.1
    w = torch.softmax(rn(B, Lq, 8, 16, seed=1), -1).view(B, Lq, 8, 4, 4)
    shp = torch.tensor(shapes, device="cuda", dtype=torch.long)
    start = torch.cat([shp.new_zeros(1), (shp[:, 0] * shp[:, 1]).cumsum(0)[:-1]])
    ref = msda.msda_core(value, shapes, loc, w)
    assert (ops.msda_forward(value, shp, start, loc, w) - ref).abs().max() < 1e-5
    vb = value.bfloat16()
    refb = msda.msda_core(vb.float(), shapes, loc, w)
    assert rel(ops.msda_forward(vb, shp, start, loc, w, out_dtype=torch.bfloat16), refb) < BF16_TOL


@pytest.mark.parametrize("oh,ow", [(200, 300), (128, 128), (130, 250), (512, 512)])
def test_blend_tiles_bit_exact(ops, oh, ow):
    from oracle import tiles
    n_h, n_w, _, _ = tiles.tile_grid(oh, ow)
    g = torch.Generator().manual_seed(31)
    tl = [torch.rand((1, 3, 512, 512), generator=g) for _ in range(n_h * n_w)]
    ref = tiles.merge_tiles(tl, (oh, ow))
    out = ops.blend_tiles(torch.cat(tl).cuda(), n_h, n_w, 64, 4 * oh, 4 * ow)
    assert torch.equal(out.cpu(), ref)


def test_blend_tiles_fixture_and_ragged_list(ops, golden):
    from oracle import tiles
    g = golden("merge_case.npz")
    oh, ow, n = (int(v) for v in g["a_size"])
    gen = torch.Generator().manual_seed(31)
    tl = torch.cat([torch.rand((1, 3, 512, 512), generator=gen) for _ in range(n)]).cuda()
    out = ops.blend_tiles(tl, 2, 3, 64, 4 * oh, 4 * ow)
    assert np.array_equal(out[0, :, ::37, ::41].cpu().numpy(), g["a_sample"])
    # fewer tiles than grid cells: the reference stops placing tiles and divides by clamp(weight, 1e-8)
    short = [t[None] for t in tl[:4].cpu()]
    ref = tiles.merge_tiles(short, (oh, ow))
    assert torch.equal(ops.blend_tiles(tl[:4].contiguous(), 2, 3, 64, 4 * oh, 4 * ow).cpu(), ref)


def test_conv_and_attention_full_size_properties(ops):
    """The two dominant kernels at the batch the bench runs (B=16).  3x3 conv 64^2 320->320: linear in its input, a shifted
    image gives the shifted output away from the border (translation equivariance of the implicit-GEMM addressing), and
    image 7 of the batch equals image 7 alone bit for bit.  Self-attention 4096^2, 5 heads: identical value rows are
    reproduced, permuting the keys with their values changes nothing beyond bf16, and every (batch, head) is independent."""
    B, H, C = 16, 64, 320
    x1, x2 = rn(B, H, H, C, seed=41).bfloat16(), rn(B, H, H, C, seed=42).bfloat16()
    w = rn(C, 9 * C, scale=(9 * C) ** -0.5, seed=43).bfloat16()
    o1, o2 = ops.conv3x3(x1, w).float(), ops.conv3x3(x2, w).float()
    o12 = ops.conv3x3((x1.float() + x2.float()).bfloat16(), w).float()
    assert rel(o12, o1 + o2) < 2e-2
    sh = torch.roll(x1, shifts=(3, 5), dims=(1, 2))
    osh = ops.conv3x3(sh.contiguous(), w).float()
    assert torch.equal(osh[:, 5:-2, 7:-2], torch.roll(o1, shifts=(3, 5), dims=(1, 2))[:, 5:-2, 7:-2])
    assert torch.equal(ops.conv3x3(x1[7:8].contiguous(), w).float(), o1[7:8])
    Hh, L = 5, 4096
    E = Hh * 64
    q, k, v = rn(B * L, E, scale=1.5, seed=44).bfloat16(), rn(B * L, E, scale=1.5, seed=45).bfloat16(), rn(B * L, E, seed=46).bfloat16()
    out = ops.attention(q, k, v, B=B, H=Hh, Lq=L, Lk=L)
    perm = torch.randperm(L, device="cuda", generator=torch.Generator(device="cuda").manual_seed(47))
    kp = k.view(B, L, E)[:, perm].reshape(B * L, E).contiguous()
    vp = v.view(B, L, E)[:, perm].reshape(B * L, E).contiguous()
    assert rel(ops.attention(q, kp, vp, B=B, H=Hh, Lq=L, Lk=L), out) < BF16_TOL
    vc = rn(1, E, seed=48).bfloat16().expand(B * L, E).contiguous()
    assert rel(ops.attention(q, k, vc, B=B, H=Hh, Lq=L, Lk=L), vc) < BF16_TOL
    one = ops.attention(q.view(B, L, E)[3].contiguous(), k.view(B, L, E)[3].contiguous(), v.view(B, L, E)[3].contiguous(),
                        B=1, H=Hh, Lq=L, Lk=L)
    assert torch.equal(one, out.view(B, L, E)[3])


def test_norms_full_size_invariances(ops):
    """Normalisation properties at the sizes of a B=16 step: GroupNorm(a x + b_g) == GroupNorm(x) for a positive scale and
    a per-group shift, LayerNorm(a x + b) == LayerNorm(x) for a per-row shift (both up to the bf16 rounding of the
    transformed input), and every image of the batch equals that image normalised alone (bit for bit)."""
    B, HW, C = 16, 4096, 320
    x = (rn(B, HW, C, seed=21) * 1.5).bfloat16()
    g, b = 1 + 0.1 * rn(C, seed=22), 0.1 * rn(C, seed=23)
    y = ops.groupnorm(x, g, b, act=ops.ACT_SILU)
    shift = rn(B, 1, 32, 1, seed=24).expand(B, HW, 32, C // 32).reshape(B, HW, C)
    y2 = ops.groupnorm((x.float() * 2.0 + shift).bfloat16(), g, b, act=ops.ACT_SILU)
    assert rel(y2, y) < 2e-2
    assert torch.equal(y[5:6], ops.groupnorm(x[5:6].contiguous(), g, b, act=ops.ACT_SILU))
    M, Cl = 151552, 256
    r = rn(M, Cl, seed=25).bfloat16()
    gl, bl = 1 + 0.1 * rn(Cl, seed=26), 0.1 * rn(Cl, seed=27)
    z = ops.layernorm(r, gl, bl)
    z2 = ops.layernorm((r.float() * 4.0 + rn(M, 1, seed=28)).bfloat16(), gl, bl)
    assert rel(z2, z) < 2e-2
    assert torch.equal(z[1000:1003], ops.layernorm(r[1000:1003].contiguous(), gl, bl))


def test_msda_fused_full_size_is_linear_in_the_values(ops):
    """Deformable attention at the TESTR encoder size (B=16, 9472 queries, 8 heads, 4 levels x 4 points) is linear in the
    value tensor for fixed sampling offsets and attention logits: msda(v1 + v2) == msda(v1) + msda(v2) up to bf16."""
    B, M, D = 16, 8, 32
    shapes = [(16, 16), (32, 32), (64, 64), (64, 64)]
    S = sum(h * w for h, w in shapes)
    shp = torch.tensor(shapes, device="cuda", dtype=torch.long)
    start = torch.cat([shp.new_zeros(1), (shp[:, 0] * shp[:, 1]).cumsum(0)[:-1]])
    v1, v2 = rn(B, S, M, D, seed=31).bfloat16(), rn(B, S, M, D, seed=32).bfloat16()
    proj = torch.cat([rn(B * S, M * 16 * 2, seed=33) * 2.0, rn(B * S, M * 16, seed=34)], 1).bfloat16().contiguous()
    refs = []
    for (h, w) in shapes:
        ry, rx = torch.meshgrid(torch.linspace(0.5, h - 0.5, h, device="cuda"), torch.linspace(0.5, w - 0.5, w, device="cuda"), indexing="ij")
        refs.append(torch.stack((rx.reshape(-1) / w, ry.reshape(-1) / h), -1))
    ref = torch.cat(refs, 0)[:, None, :].expand(S, 4, 2).contiguous()
    kw = dict(B=B, Lq=S, n_heads=M, n_levels=4, n_points=4, q_per_ref=1, ref_shared=True)
    o1 = ops.msda_fused(v1, shp, start, proj, ref, **kw).float()
    o2 = ops.msda_fused(v2, shp, start, proj, ref, **kw).float()
    o12 = ops.msda_fused((v1.float() + v2.float()).bfloat16(), shp, start, proj, ref, **kw).float()
    assert rel(o12, o1 + o2) < 2e-2
    assert torch.equal(ops.msda_fused(v1, shp, start, proj, ref, **kw).float(), o1)


def test_blend_of_constant_tiles_is_constant_at_4k(ops):
    """configs[4] geometry (3840 x 2160 LQ image, 35 x 20 = 700 tiles of 512^2, 15360 x 8640 output): blending tiles that
    all hold the same value per channel gives exactly... that value up to one rounding of the weight normalisation, on
    every output pixel (no uncovered pixel, no double-counted seam)."""
    from oracle import tiles
    oh, ow = 2160, 3840
    n_h, n_w, _, _ = tiles.tile_grid(oh, ow)
    assert n_h * n_w == 700
    vals = torch.tensor([0.25, 0.5, 0.8125], device="cuda").view(1, 3, 1, 1)
    tl = vals.expand(700, 3, 512, 512).contiguous()
    out = ops.blend_tiles(tl, n_h, n_w, 64, 4 * oh, 4 * ow)
    assert tuple(out.shape) == (1, 3, 4 * oh, 4 * ow)
    assert (out - vals).abs().max().item() < 1e-6


@pytest.mark.parametrize("M,C,N,act", [(300, 320, 960, "none"), (1000, 640, 640, "none"), (257, 64, 512, "geglu"),
                                       (4096, 1280, 10240, "geglu")])
def test_layernorm_folded_into_gemm(cuda_lib, M, C, N, act):
    """Linear(LayerNorm(x)) from the raw rows: tair_row_stats + the ln_row_stats / ln_col_sum epilogue against fp32 torch,
    with a row offset far from zero (mean >> std is where the folded form could cancel badly)."""
    import torch.nn.functional as F
    from tair_b200 import ops
    from tair_b200.model.attention import fold_layernorm, interleave_geglu
    g = torch.Generator(device="cuda").manual_seed(3)
    x = (torch.randn(M, C, device="cuda", generator=g) * 1.5 + 3.0 * torch.randn(M, 1, device="cuda", generator=g)).bfloat16()
    ln = torch.nn.LayerNorm(C).cuda()
    with torch.no_grad():
        ln.weight.copy_(1 + 0.3 * torch.randn(C, device="cuda", generator=g))
        ln.bias.copy_(0.2 * torch.randn(C, device="cuda", generator=g))
    w = torch.randn(N, C, device="cuda", generator=g) / C ** 0.5
    b = 0.1 * torch.randn(N, device="cuda", generator=g)
    ln._stamp = lambda: 0
    wf, bias, cs = fold_layernorm(w, b, ln)
    stats = ops.row_stats(x, ln.eps)
    xf = x.float()
    assert torch.allclose(stats[:, 0], xf.mean(1), atol=1e-5, rtol=1e-5)
    assert torch.allclose(stats[:, 1], (xf.var(1, unbiased=False) + ln.eps).rsqrt(), rtol=1e-4)
    y_ref = F.linear(F.layer_norm(xf, (C,), ln.weight, ln.bias, ln.eps), w, b)
    if act == "geglu":
        wi, bi = interleave_geglu(wf, bias)
        out = ops.gemm(x, wi, bias=bi, act=ops.ACT_GEGLU, ln=(stats, wi.float().sum(1).contiguous()))
        v, gate = y_ref.chunk(2, dim=-1)
        y_ref = v * F.gelu(gate)
    else:
        out = ops.gemm(x, wf, bias=bias, ln=(stats, cs))
    assert rel(out, y_ref) < 1.5e-2
    # and it agrees with the unfused kernels (LayerNorm kernel -> GEMM) to bf16 round-off
    y_unf = ops.layernorm(x, ln.weight.detach().float().contiguous(), ln.bias.detach().float().contiguous(), eps=ln.eps)
    if act != "geglu":
        assert rel(out, ops.gemm(y_unf, w.bfloat16().contiguous(), bias=b.contiguous())) < 1.5e-2


def test_gelu_epilogue_matches_erf_gelu(cuda_lib):
    """The A&S erf used by the GELU / GEGLU epilogues is exact to fp32 round-off over the whole range."""
    import torch.nn.functional as F
    from tair_b200 import ops
    x = torch.linspace(-9, 9, 4096 * 64, device="cuda").view(4096, 64).bfloat16()
    eye = torch.eye(64, device="cuda").bfloat16()
    out = ops.gemm(x, eye, act=ops.ACT_GELU, out_dtype=torch.float32)
    assert (out - F.gelu(x.float())).abs().max().item() < 2e-6


def test_geglu_gate_activation_tolerance(cuda_lib):
    """The GEGLU epilogue evaluates the gate's GELU in tanh form through the hardware tanh (csrc/common.cuh gelu_tanh_f):
    absolute deviation from the exact erf GELU <= 1e-3 everywhere (4.8e-4 analytic + tanh.approx), i.e. well below the bf16
    rounding of the product it feeds."""
    import torch.nn.functional as F
    from tair_b200 import ops
    from tair_b200.model.attention import interleave_geglu
    n = 128
    gate = torch.linspace(-8, 8, 2048 * n, device="cuda").view(2048, n)
    # value weights = 0 with bias 1 (value == 1), gate weights = identity on the first n inputs
    a = gate.bfloat16()
    w = torch.zeros(2 * n, n, device="cuda")
    w[n:] = torch.eye(n, device="cuda")
    b = torch.cat([torch.ones(n, device="cuda"), torch.zeros(n, device="cuda")])
    wi, bi = interleave_geglu(w.bfloat16(), b)
    out = ops.gemm(a, wi, bias=bi, act=ops.ACT_GEGLU, out_dtype=torch.float32)
    ref = F.gelu(a.float())
    assert (out - ref).abs().max().item() < 1e-3
    out_bf = ops.gemm(a, wi, bias=bi, act=ops.ACT_GEGLU)
    assert (out_bf.float() - ref).abs().max().item() < 1e-3 + 2 ** -8 * ref.abs().max().item()
