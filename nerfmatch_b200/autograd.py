"""torch.autograd bridges for the differentiable path.

The reference's iNeRF pose refinement (nerfmatch/nerfmatch_evaluator.py:288-500) reaches three pieces of the renderer
directly and differentiates through them w.r.t. the camera pose: ``renderer.xyz_encoder`` / ``dirs_encoder``,
``renderer.nerf_fine.forward`` and ``volume_render_radiance_field``.  Each ``Function`` below pairs the forward CUDA stage
with its input-gradient stage (include/nmr.h: nmr_stage_*_backward), so that loop runs unchanged on the shim's modules.
PyTorch only carries the graph; the one piece of arithmetic here is the two-line chain rule of ``disp_map`` onto the
depth / acc gradients.
"""
from __future__ import annotations

import torch


def wants_grad(*tensors) -> bool:
    return torch.is_grad_enabled() and any(isinstance(t, torch.Tensor) and t.requires_grad for t in tensors)


class EncodeFn(torch.autograd.Function):
    """x [P,3] (+ constant var [P,3]) -> encoding [P,dim]; backward: nmr_stage_encode_backward."""

    @staticmethod
    def forward(ctx, x, var, eng, kind):
        xf = x.detach().float().contiguous()
        vf = None if var is None else var.detach().float().contiguous()
        ctx.eng, ctx.kind = eng, kind
        ctx.save_for_backward(xf, vf)
        return eng.stage_encode(kind, xf, vf)

    @staticmethod
    def backward(ctx, g):
        xf, vf = ctx.saved_tensors
        return ctx.eng.stage_encode_backward(ctx.kind, xf, vf, g.float().contiguous()), None, None, None


class MlpFn(torch.autograd.Function):
    """NeRF.forward on x [P,in_dim] -> (raw [P,4], feat [P,256]); backward: nmr_stage_mlp_backward.  On the tensor-core
    precisions the forward kernel emits the nine layers' ReLU masks (one bit per activation) and the backward is the dgrad
    chain alone; on the fp32 path the activations are recomputed in the backward."""

    @staticmethod
    def forward(ctx, x, eng, which, stop_layer):
        xf = x.detach().float().contiguous()
        ctx.eng, ctx.which, ctx.stop_layer = eng, which, stop_layer
        ctx.set_materialize_grads(False)
        if eng.cfg.precision in ("bf16", "fp16"):
            raw, feat, masks = eng.stage_mlp_masks(which, xf, stop_layer)
            ctx.with_masks = True
            ctx.save_for_backward(raw, masks)
        else:
            raw, feat = eng.stage_mlp(which, xf, stop_layer, want_feat=True)
            ctx.with_masks = False
            ctx.save_for_backward(xf)
        return raw, feat

    @staticmethod
    def backward(ctx, g_raw, g_feat):
        if g_raw is None and g_feat is None:
            return None, None, None, None
        saved = ctx.saved_tensors
        P, dev = saved[0].shape[0], saved[0].device
        g_raw = torch.zeros(P, 4, device=dev) if g_raw is None else g_raw.float().contiguous()
        g_feat = None if g_feat is None else g_feat.float().contiguous()
        if ctx.with_masks:
            raw, masks = saved
            return ctx.eng.stage_mlp_backward(ctx.which, None, ctx.stop_layer, g_raw, g_feat, raw_fwd=raw, masks=masks), None, None, None
        return ctx.eng.stage_mlp_backward(ctx.which, saved[0], ctx.stop_layer, g_raw, g_feat), None, None, None


class CompositeFn(torch.autograd.Function):
    """volume_render_radiance_field: (raw [n,S,4], rays_d [n,3]; z constant) -> rgb, disp, acc, weights, depth;
    backward: nmr_stage_composite_backward."""

    @staticmethod
    def forward(ctx, raw, rays_d, z, eng, flags):
        rf, df, zf = raw.detach().float().contiguous(), rays_d.detach().float().contiguous(), z.detach().float().contiguous()
        out = eng.stage_composite(rf, zf, df, flags=flags)
        ctx.eng, ctx.flags = eng, flags
        ctx.need_d = rays_d.requires_grad
        ctx.save_for_backward(rf, df, zf, out["depth"], out["acc"])
        ctx.set_materialize_grads(False)
        return out["rgb"], out["disp"], out["acc"], out["weights"], out["depth"]

    @staticmethod
    def backward(ctx, g_rgb, g_disp, g_acc, g_weights, g_depth):
        rf, df, zf, depth, acc = ctx.saved_tensors
        c = lambda t: None if t is None else t.float().contiguous()
        g_depth, g_acc = c(g_depth), c(g_acc)
        if g_disp is not None:
            # disp = 1 / max(1e-10, depth / acc) (render_utils.py:222): chain it onto the depth / acc gradients
            q = depth / acc
            dq = torch.where(q > 1e-10, -g_disp.float() / (q * q), torch.zeros_like(q))
            dq = torch.nan_to_num(dq, nan=0.0, posinf=0.0, neginf=0.0)
            gd, ga = dq / acc, -dq * depth / (acc * acc)
            g_depth = gd if g_depth is None else g_depth + gd
            g_acc = ga if g_acc is None else g_acc + ga
        g_raw, g_d = ctx.eng.stage_composite_backward(rf, zf, df, flags=ctx.flags, grad_rgb=c(g_rgb), grad_depth=g_depth, grad_acc=g_acc,
                                                      grad_weights=c(g_weights), want_rays_d=ctx.need_d)
        return g_raw, g_d, None, None, None
