"""GEMM -> reduce-scatter fusion, kernel-level check (run under torchrun, >= 2 GPUs): every rank holds a locally
accumulated gradient and adds one more weight-gradient GEMM with push mode on; afterwards each owner's shard must
equal the sum over ranks of (accumulated + tile)."""
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bert_pytorch_b200.ops import api as K  # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", init_method="env://", device_id=dev)
    import torch.distributed._symmetric_memory as symm
    out = {"world": world}
    ext = K.extension()
    for name, (n_out, k_out, tokens, lead) in {"ffn": (4096, 1024, 12288, 4096), "qkv": (3072, 1024, 12288, 2048 * 3 + 512),
                                                "small": (768, 256, 512, 1000 * 8)}.items():
        numel = (lead + n_out * k_out + 4096 + 2047) // 2048 * 2048
        g = symm.empty(numel, dtype=torch.float32, device=dev)
        h = symm.rendezvous(g, dist.group.WORLD)
        per = (numel + world - 1) // world
        per = (per + 2047) // 2048 * 2048
        ext.set_grad_peers(world, rank, list(h.buffer_ptrs), numel, per)
        gen = torch.Generator(device=dev).manual_seed(100 + rank)
        acc = torch.randn(numel, device=dev, generator=gen)
        g.copy_(acc)
        dy = (torch.randn(tokens, n_out, device=dev, generator=gen) * 0.1).bfloat16()
        x = (torch.randn(tokens, k_out, device=dev, generator=gen) * 0.1).bfloat16()
        view = g[lead:lead + n_out * k_out].view(n_out, k_out)
        torch.cuda.synchronize(); dist.barrier()
        ext.set_grad_push(True)
        K.wgrad_accumulate(dy, x, view, push=True)
        ext.set_grad_push(False)
        torch.cuda.synchronize(); dist.barrier()
        expect = acc.clone()
        expect[lead:lead + n_out * k_out] += (dy.float().t() @ x.float()).reshape(-1)
        # only the GEMM region is reduced by the pushes; outside it every rank keeps its own values
        region = torch.zeros(numel, device=dev)
        region[lead:lead + n_out * k_out] = expect[lead:lead + n_out * k_out]
        dist.all_reduce(region)
        lo, hi = min(rank * per, numel), min((rank + 1) * per, numel)
        a, b = max(lo, lead), min(hi, lead + n_out * k_out)
        rec = {"owned": [a, b]}
        if b > a:
            got, want = g[a:b], region[a:b]
            err = (got - want).abs()
            tol = 1e-3 * want.abs().max().item()
            bad = (err > tol).nonzero().flatten()
            rec.update(max_err=err.max().item(), tol=tol, n_bad=int(bad.numel()), n=int(b - a))
            if bad.numel():
                idx = (bad[:8] + a - lead).tolist()
                rec["bad_rc"] = [(i // k_out, i % k_out) for i in idx]
                rec["bad_got_want_acc"] = [(float(got[j]), float(want[j]), float(acc[a + j])) for j in bad[:4].tolist()]
        allrec = [None] * world
        dist.all_gather_object(allrec, rec)
        out[name] = allrec
        ext.set_grad_peers(0, 0, [], 0, 1)
        del g, h
    if rank == 0:
        print(json.dumps(out), flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
