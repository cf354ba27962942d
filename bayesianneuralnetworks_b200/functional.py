"""torch.autograd glue around the C ABI: each Function forwards to the CUDA library on the current
stream and implements the closed-form backward of SURVEY §3.2 with the fused kernels (eps is
regenerated from the Philox counters, never stored).  No CPU path: non-CUDA tensors raise.
"""
import math

import torch

from . import _C


class DrawSpec:
    """Which eps a launch uses: Philox (seed, tensor_id, [sample_begin, sample_begin+S)) or an
    injected tensor [S, numel] (tests).  `step_dev` (optional int64[1] device tensor) is added to the
    Philox step word inside the kernels (CUDA-graph replays, runtime.graph_safe_rng)."""
    __slots__ = ("seed", "tensor_id", "sample_begin", "step", "eps", "step_dev", "signs")

    def __init__(self, seed, tensor_id, draw_begin, eps=None, step_dev=None, signs=None):
        self.seed = seed
        self.tensor_id = tensor_id
        self.sample_begin = draw_begin & 0xFFFFFFFF
        self.step = draw_begin >> 32
        self.eps = eps
        self.step_dev = step_dev
        self.signs = signs            # (R [S, out], S [S, in]): rank-one sign noise instead of Philox normals (Flipout)

    def rng(self, elem_offset=0):
        return _C.make_rng(self.seed, self.step, self.tensor_id, elem_offset=elem_offset, step_dev=self.step_dev,
                           signs=self.signs, sample_begin=self.sample_begin)


def _check_f32_cuda(name, t):
    if t is None:
        return
    _C.require_cuda(t)
    if t.dtype != torch.float32:
        raise TypeError(f"{name}: the variational hot path computes in float32, got {t.dtype}")


def _zero_grads(w_shape, n_bias, device):
    """Zeroed accumulators [2, *w_shape] (d mean, d scale of the weight) and [2, n_bias] (of the bias) carved out of ONE
    buffer: the gradient kernels accumulate into them, so a layer's backward needs a single fill."""
    nw = 2 * math.prod(w_shape) if w_shape is not None else 0
    nb = 2 * n_bias if n_bias is not None else 0
    if nw + nb == 0:
        return None, None
    flat = torch.zeros(nw + nb, device=device, dtype=torch.float32)
    grads = flat[:nw].view((2,) + tuple(w_shape)) if nw else None
    bg = flat[nw:].view(2, n_bias) if nb else None
    return grads, bg


def _eps_slice(spec, S, numel, lo=None, hi=None):
    """Injected eps as a contiguous [S, n] block (optionally the [lo, hi) slice of every sample)."""
    if spec is None or spec.eps is None:
        return None
    e = spec.eps.reshape(S, numel)
    if lo is not None:
        e = e[:, lo:hi]
    return e.contiguous()


# ------------------------------------------------------------------------------------------------
class SampledLinear(torch.autograd.Function):
    """y[s] = x[s] W_s^T + b_s for S Monte-Carlo samples in one launch (dense.py:46-60).

    x: [R, K] with R = S*M rows (sample-major) or, when `shared`, [M, K] used by every sample.
    Returns [S*M, N]."""

    @staticmethod
    def forward(ctx, x, mu_w, rho_w, mu_b, rho_b, S, shared, spec_w, spec_b, precision):
        for n, t in (("input", x), ("weight.mean", mu_w), ("weight.scale", rho_w), ("bias.mean", mu_b),
                     ("bias.scale", rho_b)):
            _check_f32_cuda(n, t)
        x = x.contiguous()
        N, K = mu_w.shape
        if x.dim() != 2 or x.shape[1] != K:
            raise RuntimeError(f"sampled linear: input {tuple(x.shape)} does not match in_features {K}")
        rows = x.shape[0]
        M = rows if shared else rows // S
        if not shared and M * S != rows:
            raise RuntimeError(f"sampled linear: {rows} rows are not divisible by {S} Monte-Carlo samples")
        mu_w_c, rho_w_c = mu_w.contiguous(), rho_w.contiguous()
        sigma_w = _C.stddev(rho_w_c)
        has_bias = mu_b is not None
        sigma_b = _C.stddev(rho_b.contiguous()) if has_bias else None
        y = torch.empty((S * M, N), device=x.device, dtype=torch.float32)
        _C.sampled_gemm_fwd(x, K, 0 if shared else M * K, mu_w_c, sigma_w,
                            mu_b.contiguous() if has_bias else None, sigma_b,
                            _eps_slice(spec_w, S, N * K), _eps_slice(spec_b, S, N) if has_bias else None,
                            _C.make_view(y.data_ptr(), N, 1), M * N, M, N, K, S, spec_w.sample_begin,
                            spec_w.rng(), spec_b.rng() if has_bias else None, precision)
        ctx.save_for_backward(x, mu_w_c, rho_w_c, sigma_w, rho_b)
        ctx.meta = (S, shared, spec_w, spec_b, precision, M, N, K)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, mu_w, rho_w, sigma_w, rho_b = ctx.saved_tensors
        S, shared, spec_w, spec_b, precision, M, N, K = ctx.meta
        ldy = N
        if precision == _C.PREC_TF32 and N % 4 != 0:
            # the TMA-fed kernels need a 16-byte row pitch: pad the (narrow) gradient rows, e.g. a 10-class head
            ldy = (N + 3) // 4 * 4
            dy = torch.nn.functional.pad(dy, (0, ldy - N))
        dy = dy.contiguous()
        dy_view = _C.make_view(dy.data_ptr(), ldy, 1)
        eps_w = _eps_slice(spec_w, S, N * K)
        a_stride = 0 if shared else M * K
        dx = dmu_w = drho_w = dmu_b = drho_b = None
        if ctx.needs_input_grad[0]:
            dx = torch.empty_like(x)
            _C.sampled_gemm_dgrad(dy_view, M * ldy, mu_w, sigma_w, eps_w, dx, K, a_stride, M, N, K, S,
                                  spec_w.sample_begin, spec_w.rng(), precision)
        need_w = ctx.needs_input_grad[1] or ctx.needs_input_grad[2]
        need_b = rho_b is not None and (ctx.needs_input_grad[3] or ctx.needs_input_grad[4])
        grads, bg = _zero_grads((N, K) if need_w else None, N if need_b else None, x.device)
        if need_w:
            dmu_w, drho_w = grads[0], grads[1]
            _C.sampled_gemm_wgrad(dy_view, M * ldy, x, K, a_stride, rho_w, eps_w, dmu_w, drho_w, M, N, K, S,
                                  spec_w.sample_begin, spec_w.rng(), precision)
        if need_b:
            dmu_b, drho_b = bg[0], bg[1]
            _C.bias_grad(dy_view, M * ldy, rho_b.contiguous(), _eps_slice(spec_b, S, N), dmu_b, drho_b, M, N, S,
                         spec_b.sample_begin, spec_b.rng())
        return dx, dmu_w, drho_w, dmu_b, drho_b, None, None, None, None, None


# ------------------------------------------------------------------------------------------------
def _conv_out(size, k, s, p, d):
    return (size + 2 * p - d * (k - 1) - 1) // s + 1


_KEEP_COL_BYTES = 2 << 30      # im2col matrices up to this size are kept from forward to backward


class SampledConv2d(torch.autograd.Function):
    """y[s] = conv2d(x[s], W_s, b_s) for S samples (conv.py:65-73,112-119): im2col lowering of the
    activations per group, then the sampled GEMM writing straight into the NCHW output.

    x: [S*B, C, H, W] (sample-major) or, when `shared`, [B, C, H, W].  Returns [S*B, Cout, OH, OW]."""

    @staticmethod
    def forward(ctx, x, mu_w, rho_w, mu_b, rho_b, S, shared, spec_w, spec_b, precision, stride, padding,
                dilation, groups):
        for n, t in (("input", x), ("weight.mean", mu_w), ("weight.scale", rho_w), ("bias.mean", mu_b),
                     ("bias.scale", rho_b)):
            _check_f32_cuda(n, t)
        x = x.contiguous()
        Cout, Cg, KH, KW = mu_w.shape
        if x.dim() != 4 or x.shape[1] != Cg * groups:
            raise RuntimeError(f"sampled conv2d: input {tuple(x.shape)} does not match weight {tuple(mu_w.shape)} "
                               f"with groups={groups}")
        rows, C, H, W = x.shape
        B = rows if shared else rows // S
        if not shared and B * S != rows:
            raise RuntimeError(f"sampled conv2d: batch {rows} is not divisible by {S} Monte-Carlo samples")
        OH = _conv_out(H, KH, stride[0], padding[0], dilation[0])
        OW = _conv_out(W, KW, stride[1], padding[1], dilation[1])
        if OH <= 0 or OW <= 0:
            raise RuntimeError("sampled conv2d: empty output")
        geo = (B, C, H, W, Cout, Cg, KH, KW, OH, OW, tuple(stride), tuple(padding), tuple(dilation), groups)
        mu_w_c, rho_w_c = mu_w.contiguous(), rho_w.contiguous()
        sigma_w = _C.stddev(rho_w_c)
        has_bias = mu_b is not None
        mu_b_c = mu_b.contiguous() if has_bias else None
        sigma_b = _C.stddev(rho_b.contiguous()) if has_bias else None
        y = torch.empty((S * B, Cout, OH, OW), device=x.device, dtype=torch.float32)
        Ng, Kg, P = Cout // groups, Cg * KH * KW, OH * OW
        M = B * P
        col = torch.empty((rows * P, Kg), device=x.device, dtype=torch.float32)
        for g in range(groups):
            _C.im2col(x, col, SampledConv2d._geom(rows, geo, g))
            lo, hi = g * Ng * Kg, (g + 1) * Ng * Kg
            view = _C.make_view(y.data_ptr() + 4 * g * Ng * P, Cout * P, P)
            _C.sampled_gemm_fwd(col, Kg, 0 if shared else M * Kg, mu_w_c.view(-1)[lo:hi], sigma_w.view(-1)[lo:hi],
                                mu_b_c[g * Ng:(g + 1) * Ng] if has_bias else None,
                                sigma_b[g * Ng:(g + 1) * Ng] if has_bias else None,
                                _eps_slice(spec_w, S, Cout * Kg, lo, hi),
                                _eps_slice(spec_b, S, Cout, g * Ng, (g + 1) * Ng) if has_bias else None,
                                view, B * Cout * P, M, Ng, Kg, S, spec_w.sample_begin, spec_w.rng(lo),
                                spec_b.rng(g * Ng) if has_bias else None, precision)
        ctx.save_for_backward(x, mu_w_c, rho_w_c, sigma_w, rho_b)
        ctx.meta = (S, shared, spec_w, spec_b, precision, geo)
        # the weight gradient contracts dY with the same im2col matrix: keep it for the backward pass instead of
        # lowering x a second time (one launch and one write of the matrix less), unless it is very large
        keep = groups == 1 and (ctx.needs_input_grad[1] or ctx.needs_input_grad[2]) and col.numel() * 4 <= _KEEP_COL_BYTES
        ctx.col = col if keep else None
        return y

    @staticmethod
    def _geom(batch, geo, g):
        B, C, H, W, Cout, Cg, KH, KW, OH, OW, stride, padding, dilation, groups = geo
        return _C.bnn_conv2d_geom(batch, C, H, W, g * Cg, Cg, KH, KW, OH, OW, stride[0], stride[1],
                                  padding[0], padding[1], dilation[0], dilation[1])

    @staticmethod
    def backward(ctx, dy):
        x, mu_w, rho_w, sigma_w, rho_b = ctx.saved_tensors
        S, shared, spec_w, spec_b, precision, geo = ctx.meta
        B, C, H, W, Cout, Cg, KH, KW, OH, OW, stride, padding, dilation, groups = geo
        dy = dy.contiguous()
        rows = x.shape[0]
        Ng, Kg, P = Cout // groups, Cg * KH * KW, OH * OW
        M = B * P
        a_stride = 0 if shared else M * Kg
        # TF32: hand dY to the kernels as a row-major [S*B*P, Cout] matrix (one strided copy), which makes the
        # TMA-fed data- and weight-gradient kernels eligible; fp32 mode reads the NCHW tensor in place.
        rowmajor_dy = precision == _C.PREC_TF32 and P > 1 and Cout % 4 == 0 and Ng % 4 == 0
        if rowmajor_dy:
            dy_rows = dy.view(S * B, Cout, P).transpose(1, 2).contiguous()
        need_x = ctx.needs_input_grad[0]
        need_w = ctx.needs_input_grad[1] or ctx.needs_input_grad[2]
        need_b = rho_b is not None and (ctx.needs_input_grad[3] or ctx.needs_input_grad[4])
        dx = torch.empty_like(x) if need_x else None
        grads, bg = _zero_grads(tuple(mu_w.shape) if need_w else None, Cout if need_b else None, x.device)
        kept = ctx.col if need_w else None
        ctx.col = None
        col = kept if kept is not None else (
            torch.empty((rows * P, Kg), device=x.device, dtype=torch.float32) if need_w else None)
        dcol = torch.empty((rows * P, Kg), device=x.device, dtype=torch.float32) if need_x else None
        for g in range(groups):
            lo, hi = g * Ng * Kg, (g + 1) * Ng * Kg
            if rowmajor_dy:
                dy_view, dy_ss = _C.make_view(dy_rows.data_ptr() + 4 * g * Ng, Cout, 1), M * Cout
            else:
                dy_view, dy_ss = _C.make_view(dy.data_ptr() + 4 * g * Ng * P, Cout * P, P), B * Cout * P
            eps_w = _eps_slice(spec_w, S, Cout * Kg, lo, hi)
            geom = SampledConv2d._geom(rows, geo, g)
            if need_x:
                _C.sampled_gemm_dgrad(dy_view, dy_ss, mu_w.view(-1)[lo:hi], sigma_w.view(-1)[lo:hi], eps_w,
                                      dcol, Kg, a_stride, M, Ng, Kg, S, spec_w.sample_begin, spec_w.rng(lo),
                                      precision)
                _C.col2im(dcol, dx, geom, False)
            if need_w:
                if kept is None:
                    _C.im2col(x, col, geom)
                _C.sampled_gemm_wgrad(dy_view, dy_ss, col, Kg, a_stride, rho_w.view(-1)[lo:hi], eps_w,
                                      grads[0].view(-1)[lo:hi], grads[1].view(-1)[lo:hi], M, Ng, Kg, S,
                                      spec_w.sample_begin, spec_w.rng(lo), precision)
            if need_b:
                _C.bias_grad(dy_view, dy_ss, rho_b.contiguous()[g * Ng:(g + 1) * Ng],
                             _eps_slice(spec_b, S, Cout, g * Ng, (g + 1) * Ng), bg[0][g * Ng:(g + 1) * Ng],
                             bg[1][g * Ng:(g + 1) * Ng], M, Ng, S, spec_b.sample_begin, spec_b.rng(g * Ng))
        return (dx, grads[0] if need_w else None, grads[1] if need_w else None,
                bg[0] if need_b else None, bg[1] if need_b else None) + (None,) * 9


# ------------------------------------------------------------------------------------------------
def conv_implicit_eligible(in_channels, groups):
    """Layers whose eps stream is keyed in (o, kh, kw, c) order and that run the implicit-GEMM path (a STATIC property of
    the layer, so that forward, backward and `.sampled` agree whatever the precision mode or the input layout)."""
    return groups == 1 and in_channels % 32 == 0


_CONV_OUTPUT = {"format": "contiguous"}


def set_conv_output_format(fmt):
    """Memory format of the implicit-GEMM conv layers' OUTPUT: 'contiguous' (default: NCHW-contiguous, what the reference
    produces and what a following Flatten / view wants; the gradient that comes back is transposed to channels-last by one
    fused pass that also forms the bias gradient) or 'preserve' (torch's convention: a channels_last input gives a
    channels_last result — the choice for conv -> conv stacks kept in channels_last)."""
    if fmt not in ("contiguous", "preserve"):
        raise ValueError("conv output format must be 'contiguous' or 'preserve'")
    _CONV_OUTPUT["format"] = fmt


def _nhwc(t):
    """(tensor whose memory is NHWC-contiguous, copied?) for a logical NCHW tensor."""
    if t.permute(0, 2, 3, 1).is_contiguous():
        return t
    return t.contiguous(memory_format=torch.channels_last)


class SampledConv2dImplicit(torch.autograd.Function):
    """y[s] = conv2d(x[s], W_s, b_s) for S samples (conv.py:65-73,112-119) WITHOUT an im2col matrix: channels-last
    activations reach the tensor cores through a TMA im2col tensor map, the weights are presented in (o, kh, kw, c) order
    (one small layout kernel per step, which also produces sigma), the input gradient of stride-1 layers is the
    transposed-filter gather, the weight gradient reads dY and x through tensor maps and is returned in OIHW order.
    TF32 mode; fp32 (3xTF32) mode and the input gradient of strided layers lower explicitly in the same column order
    (bnn_im2col_nhwc / bnn_col2im_nhwc + the register-staged / TMA GEMM kernels).

    x: logical [S*B, C, H, W] (sample-major) or, when `shared`, [B, C, H, W]; any memory format — channels_last input gives
    a channels_last result under set_conv_output_format('preserve'); by default the result is NCHW-contiguous (written
    through the kernel's strided epilogue) and a contiguous NCHW input is converted once."""

    @staticmethod
    def forward(ctx, x, mu_w, rho_w, mu_b, rho_b, S, shared, spec_w, spec_b, precision, stride, padding, dilation):
        for n, t in (("input", x), ("weight.mean", mu_w), ("weight.scale", rho_w), ("bias.mean", mu_b),
                     ("bias.scale", rho_b)):
            _check_f32_cuda(n, t)
        Cout, C, KH, KW = mu_w.shape
        if x.dim() != 4 or x.shape[1] != C:
            raise RuntimeError(f"sampled conv2d: input {tuple(x.shape)} does not match weight {tuple(mu_w.shape)}")
        rows, _, H, W = x.shape
        B = rows if shared else rows // S
        if not shared and B * S != rows:
            raise RuntimeError(f"sampled conv2d: batch {rows} is not divisible by {S} Monte-Carlo samples")
        OH = _conv_out(H, KH, stride[0], padding[0], dilation[0])
        OW = _conv_out(W, KW, stride[1], padding[1], dilation[1])
        if OH <= 0 or OW <= 0:
            raise RuntimeError("sampled conv2d: empty output")
        nchw_out = _CONV_OUTPUT["format"] == "contiguous" or x.is_contiguous()
        x_cl = _nhwc(x)
        geom = _C.conv_geom(B, H, W, C, OH, OW, Cout, KH, KW, stride, padding, dilation)
        wl = _C.conv_weight_layout(mu_w.contiguous(), rho_w.contiguous())          # [3, numel]: mean, sigma, scale
        has_bias = mu_b is not None
        mu_b_c = mu_b.contiguous() if has_bias else None
        sigma_b = _C.stddev(rho_b.contiguous()) if has_bias else None
        P, K = OH * OW, KH * KW * C
        M = B * P
        y = torch.empty((S * B, Cout, OH, OW), device=x.device, dtype=torch.float32,
                        memory_format=torch.contiguous_format if nchw_out else torch.channels_last)
        if nchw_out:
            y_view, y_ss = _C.make_view(y.data_ptr(), Cout * P, P), B * Cout * P
        else:
            y_view, y_ss = _C.make_view(y.data_ptr(), Cout, 1), M * Cout
        eps_w = _eps_slice(spec_w, S, Cout * K)
        eps_b = _eps_slice(spec_b, S, Cout) if has_bias else None
        x_ss = 0 if shared else B * H * W * C
        if precision == _C.PREC_TF32:
            _C.sampled_conv2d_fwd(x_cl, x_ss, wl[0], wl[1], mu_b_c, sigma_b, eps_w, eps_b, y_view, y_ss, geom, S,
                                  spec_w.sample_begin, spec_w.rng(), spec_b.rng() if has_bias else None)
        else:
            col = _C.im2col_nhwc(x_cl, geom, rows)
            _C.sampled_gemm_fwd(col, K, 0 if shared else M * K, wl[0], wl[1], mu_b_c, sigma_b, eps_w, eps_b, y_view, y_ss,
                                M, Cout, K, S, spec_w.sample_begin, spec_w.rng(), spec_b.rng() if has_bias else None,
                                precision)
        ctx.save_for_backward(x_cl, wl, rho_b)
        ctx.meta = (S, shared, spec_w, spec_b, precision, geom, tuple(mu_w.shape), rows)
        return y

    @staticmethod
    def backward(ctx, dy):
        x_cl, wl, rho_b = ctx.saved_tensors
        S, shared, spec_w, spec_b, precision, geom, w_shape, rows = ctx.meta
        Cout, C, KH, KW = w_shape
        B, H, W, OH, OW = geom.B, geom.H, geom.W, geom.OH, geom.OW
        P, K = OH * OW, KH * KW * C
        M = B * P
        eps_w = _eps_slice(spec_w, S, Cout * K)
        x_ss = 0 if shared else B * H * W * C
        tf32 = precision == _C.PREC_TF32
        need_x = ctx.needs_input_grad[0]
        need_w = ctx.needs_input_grad[1] or ctx.needs_input_grad[2]
        need_b = rho_b is not None and (ctx.needs_input_grad[3] or ctx.needs_input_grad[4])
        # the gradient kernels read dY as rows [S*B*OH*OW, Cout]: free for a channels_last gradient; an NCHW-contiguous
        # one goes through ONE transposing pass that forms the bias gradient on the way
        bias_done = False
        bg_fused = None
        if dy.permute(0, 2, 3, 1).is_contiguous():
            dy_cl = dy
        else:
            dy_cl = None
            if dy.is_contiguous():
                if need_b:
                    bg_fused = torch.zeros((2, Cout), device=dy.device, dtype=torch.float32)
                dy_cl = _C.nchw_to_nhwc_bias_grad(dy, B, rho_b.contiguous() if need_b else None,
                                                  _eps_slice(spec_b, S, Cout) if need_b else None,
                                                  bg_fused[0] if need_b else None, bg_fused[1] if need_b else None,
                                                  spec_b.sample_begin if need_b else 0, spec_b.rng() if need_b else None)
                bias_done = dy_cl is not None and need_b
            if dy_cl is None:
                dy_cl = dy.contiguous(memory_format=torch.channels_last)
        dy_view, dy_ss = _C.make_view(dy_cl.data_ptr(), Cout, 1), M * Cout
        dx = None
        if need_x:
            dx = torch.empty(x_cl.shape, device=dy.device, dtype=torch.float32, memory_format=torch.channels_last)
            if tf32 and geom.sh == 1 and geom.sw == 1 and Cout % 32 == 0:
                _C.sampled_conv2d_dgrad(dy_cl, wl[0], wl[1], eps_w, dx, x_ss, geom, S, spec_w.sample_begin, spec_w.rng())
            else:                                       # strided layers / fp32 mode: column gradients + gather
                dcol = torch.empty((rows * P, K), device=dy.device, dtype=torch.float32)
                _C.sampled_gemm_dgrad(dy_view, dy_ss, wl[0], wl[1], eps_w, dcol, K, 0 if shared else M * K, M, Cout, K, S,
                                      spec_w.sample_begin, spec_w.rng(), precision)
                _C.col2im_nhwc(dcol, dx, geom, rows)
        grads_p, bg = _zero_grads((Cout * K,) if need_w else None, Cout if (need_b and not bias_done) else None, dy.device)
        if bias_done:
            bg = bg_fused
        dmu = drho = None
        if need_w:
            if tf32 and Cout % 4 == 0:
                _C.sampled_conv2d_wgrad(dy_cl, x_cl, x_ss, wl[2], eps_w, grads_p[0], grads_p[1], geom, S,
                                        spec_w.sample_begin, spec_w.rng())
            else:
                col = _C.im2col_nhwc(x_cl, geom, rows)
                _C.sampled_gemm_wgrad(dy_view, dy_ss, col, K, 0 if shared else M * K, wl[2], eps_w, grads_p[0], grads_p[1],
                                      M, Cout, K, S, spec_w.sample_begin, spec_w.rng(), precision)
            g = _C.conv_weight_unlayout(grads_p, w_shape)
            dmu, drho = g[0], g[1]
        if need_b and not bias_done:
            _C.bias_grad(dy_view, dy_ss, rho_b.contiguous(), _eps_slice(spec_b, S, Cout), bg[0], bg[1], M, Cout, S,
                         spec_b.sample_begin, spec_b.rng())
        return (dx, dmu, drho, bg[0] if need_b else None, bg[1] if need_b else None) + (None,) * 8


# ------------------------------------------------------------------------------------------------
class Materialize(torch.autograd.Function):
    """W_s = mean + stddev * eps_s for s in [0, S) as a tensor [S, *shape] (core.py:44-45); used by
    `.sampled` and by layers without a fused contraction (NormalConv3d)."""

    @staticmethod
    def forward(ctx, mu, rho, S, spec):
        _check_f32_cuda("mean", mu)
        _check_f32_cuda("scale", rho)
        mu_c, rho_c = mu.contiguous(), rho.contiguous()
        sigma = _C.stddev(rho_c)
        out = _C.materialize(mu_c, sigma, S, spec.sample_begin, spec.rng(), eps_in=_eps_slice(spec, S, mu.numel()))
        ctx.save_for_backward(rho_c)
        ctx.meta = (S, spec)
        return out

    @staticmethod
    def backward(ctx, g):
        (rho,) = ctx.saved_tensors
        S, spec = ctx.meta
        if spec.eps is not None:
            eps = spec.eps.reshape((S,) + tuple(rho.shape))
        else:
            zero, one = torch.zeros_like(rho), torch.ones_like(rho)
            eps = _C.materialize(zero, one, S, spec.sample_begin, spec.rng())
        dmu = g.sum(0)
        drho = (g * eps).sum(0) * torch.sigmoid(rho)
        return dmu, drho, None, None


# ------------------------------------------------------------------------------------------------
class KLSum(torch.autograd.Function):
    """sum_t coeff_t * sum_i KL(N(mu_ti, sigma_ti) || N(loc_t, scale_t)) over many tensors in one
    bandwidth-bound launch (loss.py:16-38 with coeff_t = 1 / (numel_t * n_tensors * n_batches)).
    Backward is one more pass that writes coeff_t * dKL/d(mu, rho) scaled by the upstream gradient
    (read on the device, no host sync)."""

    @staticmethod
    def forward(ctx, priors, coeffs, *params):
        n = len(params) // 2
        entries = []
        for t in range(n):
            mu, rho = params[2 * t], params[2 * t + 1]
            _check_f32_cuda("mean", mu)
            _check_f32_cuda("scale", rho)
            entries.append((mu.contiguous(), rho.contiguous(), None, None, priors[t][0], priors[t][1], coeffs[t]))
        total = _C.kl(entries, want_sums=False, want_total=True)
        ctx.save_for_backward(*[e[i] for e in entries for i in (0, 1)])
        ctx.meta = (priors, coeffs)
        return total

    @staticmethod
    def backward(ctx, g):
        priors, coeffs = ctx.meta
        saved = ctx.saved_tensors
        n = len(saved) // 2
        scale = g.detach().to(torch.float32).reshape(1).contiguous()
        entries, grads = [], []
        for t in range(n):
            mu, rho = saved[2 * t], saved[2 * t + 1]
            gm, gr = torch.empty_like(mu), torch.empty_like(rho)
            grads += [gm, gr]
            entries.append((mu, rho, gm, gr, priors[t][0], priors[t][1], coeffs[t]))
        _C.kl(entries, want_sums=False, want_total=False, grad_scale=scale)
        return (None, None) + tuple(grads)


# ------------------------------------------------------------------------------------------------
class MCCrossEntropy(torch.autograd.Function):
    """Mean cross-entropy of the [S*B, C] score matrix `x` against the B labels shared by the S Monte-Carlo samples
    (examples/MNIST/train.py:59-61 with CrossEntropyLoss): one forward and one backward kernel
    (bnn_mc_cross_entropy_fwd / _bwd) instead of torch's log_softmax + nll_loss chains over a replicated target."""

    @staticmethod
    def forward(ctx, x, target, ignore_index):
        if x.stride(1) != 1 or x.stride(0) < x.shape[1]:
            x = x.contiguous()
        target = target.contiguous()
        loss, lse, count = _C.mc_cross_entropy_fwd(x, target, ignore_index)
        ctx.save_for_backward(x, target, lse, count)
        ctx.ignore_index = ignore_index
        return loss

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, g):
        x, target, lse, count = ctx.saved_tensors
        if not ctx.needs_input_grad[0]:
            return None, None, None
        g = g.detach().to(torch.float32).reshape(()).contiguous()
        return _C.mc_cross_entropy_bwd(x, target, lse, count, g, ctx.ignore_index), None, None
