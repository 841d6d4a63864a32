"""Diagnostic (not collected by pytest): per-parameter gradient agreement of the sm_100a path with the fp32 oracle,
next to what STOCK PyTorch bf16 autocast achieves on the same step (the oracle restatement run on the GPU under
torch.autocast(bf16)) -- i.e. how much of the deviation is inherent to bf16 operands.

    python tests/grad_parity_report.py [case]      # needs a GPU
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")]

import ccd_oracle as O  # noqa: E402
from test_pretrain_parity_gpu import CASES, build, cosine, run_step  # noqa: E402


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "cfg1_tiny_b4"
    arch, E, B, K, sseed, tseed, std, norm_last = CASES[name]
    student, teacher, ssd, tsd = build(arch, E, K, sseed, tseed, std, norm_last)
    loss, loss_mod, so, to, _, (x, masks, metrics, center0) = run_step(student, teacher, K, B)

    def oracle_grads(device, autocast):
        sd = {k: v.clone().to(device).requires_grad_(v.dtype.is_floating_point and "running" not in k) for k, v in ssd.items()}
        td = {k: v.to(device) for k, v in tsd.items()}
        with torch.autocast("cuda", dtype=torch.bfloat16, enabled=autocast):
            L, _ = O.pretrain_loss(sd, td, arch, x.to(device), metrics.to(device), masks.to(device), center0.to(device), 0, 0.04)
        L.backward()
        return L.item(), {k: v.grad.detach().float().cpu() for k, v in sd.items() if v.grad is not None}

    l32, g32 = oracle_grads("cuda", False)
    lbf, gbf = oracle_grads("cuda", True)
    print(f"loss: ccd_b200 {loss.item():.6f}  oracle fp32 {l32:.6f}  stock autocast-bf16 {lbf:.6f}")
    print(f"{'parameter':52s} {'|g|':>10s} {'cos ours':>9s} {'cos stock-bf16':>14s} {'relL2 ours':>10s} {'relL2 stock':>11s}")
    worst = []
    for n, p in student.named_parameters():
        if n not in g32 or g32[n].norm() < 1e-7 or p.grad is None:
            continue
        a, r, s = p.grad.float().cpu(), g32[n], gbf[n]
        c1, c2 = cosine(a, r), cosine(s, r)
        e1 = ((a - r).norm() / r.norm()).item()
        e2 = ((s - r).norm() / r.norm()).item()
        worst.append((c1, n))
        print(f"{n:52s} {r.norm().item():10.3e} {c1:9.5f} {c2:14.5f} {e1:10.3e} {e2:11.3e}")
    worst.sort()
    print("lowest cosines (ours):", worst[:8])


if __name__ == "__main__":
    main()
