import math
import torch


class FusedLAMB(torch.optim.Optimizer):
    """apex.optimizers.FusedLAMB semantics (defaults included) on torch._foreach ops."""

    def __init__(self, params, lr=1e-3, bias_correction=True, betas=(0.9, 0.999), eps=1e-6, weight_decay=0.01,
                 amsgrad=False, adam_w_mode=True, grad_averaging=True, set_grad_none=True, max_grad_norm=1.0,
                 use_nvlamb=False):
        defaults = dict(lr=lr, bias_correction=bias_correction, betas=betas, eps=eps, weight_decay=weight_decay,
                        grad_averaging=grad_averaging, max_grad_norm=max_grad_norm)
        super().__init__(params, defaults)
        self.adam_w_mode = 1 if adam_w_mode else 0
        self.set_grad_none = set_grad_none
        self.use_nvlamb = use_nvlamb

    def zero_grad(self, set_to_none=True):
        super().zero_grad(set_to_none=self.set_grad_none)

    @torch.no_grad()
    def step(self, closure=None):
        loss = closure() if closure is not None else None
        grads_all = [p.grad for g in self.param_groups for p in g["params"] if p.grad is not None]
        if not grads_all:
            return loss
        norms = torch._foreach_norm(grads_all)
        gnorm = torch.linalg.vector_norm(torch.stack(norms))
        for group in self.param_groups:
            params = [p for p in group["params"] if p.grad is not None]
            if not params:
                continue
            grads = [p.grad for p in params]
            group["step"] = group.get("step", 0) + 1
            b1, b2 = group["betas"]
            step = group["step"]
            bc1 = 1 - b1 ** step if group["bias_correction"] else 1.0
            bc2 = 1 - b2 ** step if group["bias_correction"] else 1.0
            b3 = 1 - b1 if group["grad_averaging"] else 1.0
            mg = group["max_grad_norm"]
            clip = torch.clamp(gnorm / mg, min=1.0) if mg and mg > 0 else torch.ones((), device=gnorm.device)
            ms, vs = [], []
            for p in params:
                st = self.state[p]
                if "exp_avg" not in st:
                    st["exp_avg"] = torch.zeros_like(p)
                    st["exp_avg_sq"] = torch.zeros_like(p)
                ms.append(st["exp_avg"]); vs.append(st["exp_avg_sq"])
            g = torch._foreach_div(grads, clip)
            torch._foreach_mul_(ms, b1); torch._foreach_add_(ms, g, alpha=b3)
            torch._foreach_mul_(vs, b2); torch._foreach_addcmul_(vs, g, g, value=1 - b2)
            denom = torch._foreach_div(vs, bc2); torch._foreach_sqrt_(denom); torch._foreach_add_(denom, group["eps"])
            upd = torch._foreach_div(ms, bc1); torch._foreach_div_(upd, denom)
            wd = group["weight_decay"]
            if wd != 0:
                torch._foreach_add_(upd, params, alpha=wd)
            if wd != 0 or self.use_nvlamb:
                pn = torch._foreach_norm(params); un = torch._foreach_norm(upd)
                ratios = [torch.where((a > 0) & (b > 0), a / b, torch.ones_like(a)) * (-group["lr"]) for a, b in zip(pn, un)]
                torch._foreach_mul_(upd, ratios)
                torch._foreach_add_(params, upd)
            else:
                torch._foreach_add_(params, upd, alpha=-group["lr"])
        return loss


class FusedAdam(torch.optim.Optimizer):
    def __init__(self, params, lr=1e-3, bias_correction=True, betas=(0.9, 0.999), eps=1e-8, adam_w_mode=True,
                 weight_decay=0.0, amsgrad=False, set_grad_none=True):
        super().__init__(params, dict(lr=lr, bias_correction=bias_correction, betas=betas, eps=eps,
                                      weight_decay=weight_decay))
        self.adam_w_mode = adam_w_mode

    @torch.no_grad()
    def step(self, closure=None):
        loss = closure() if closure is not None else None
        for group in self.param_groups:
            params = [p for p in group["params"] if p.grad is not None]
            if not params:
                continue
            group["step"] = group.get("step", 0) + 1
            b1, b2 = group["betas"]; step = group["step"]
            bc1 = 1 - b1 ** step if group["bias_correction"] else 1.0
            bc2 = 1 - b2 ** step if group["bias_correction"] else 1.0
            ms, vs = [], []
            for p in params:
                st = self.state[p]
                if "exp_avg" not in st:
                    st["exp_avg"] = torch.zeros_like(p); st["exp_avg_sq"] = torch.zeros_like(p)
                ms.append(st["exp_avg"]); vs.append(st["exp_avg_sq"])
            g = [p.grad for p in params]
            torch._foreach_mul_(ms, b1); torch._foreach_add_(ms, g, alpha=1 - b1)
            torch._foreach_mul_(vs, b2); torch._foreach_addcmul_(vs, g, g, value=1 - b2)
            denom = torch._foreach_div(vs, bc2); torch._foreach_sqrt_(denom); torch._foreach_add_(denom, group["eps"])
            upd = torch._foreach_div(ms, bc1); torch._foreach_div_(upd, denom)
            if group["weight_decay"] != 0:
                torch._foreach_add_(upd, params, alpha=group["weight_decay"])
            torch._foreach_add_(params, upd, alpha=-group["lr"])
        return loss
