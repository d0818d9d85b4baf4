"""Diagnostics for the tcgen05 GEMM path: run tiny structured problems and print where
the result deviates (which K slice / row block / column block), to debug descriptor or
swizzle mistakes from a single GPU run."""
import sys
import os
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ips_b200 import ops  # noqa: E402

dev = torch.device('cuda:0')
torch.manual_seed(0)


def report(tag, got, ref):
    err = (got - ref).abs()
    rel = err.max().item() / (ref.abs().max().item() + 1e-9)
    print(f'{tag}: max abs err {err.max().item():.4e} rel {rel:.3e}')
    if rel > 1e-3:
        M, N = got.shape
        rb = err.view(M // 8 if M % 8 == 0 else M, -1, N).amax(dim=(1, 2)) if M % 8 == 0 else err.amax(1)
        print('   row-block(8) max err:', [f'{v:.2g}' for v in rb[:16].tolist()])
        cb = err.amax(0).view(-1, 8).amax(1)
        print('   col-block(8) max err:', [f'{v:.2g}' for v in cb[:16].tolist()])
        print('   got[0,:8]', got[0, :8].tolist())
        print('   ref[0,:8]', ref[0, :8].tolist())
        print('   got[1,:8]', got[1, :8].tolist())
        print('   ref[1,:8]', ref[1, :8].tolist())
    return rel


def main():
    for (M, N, K) in [(128, 64, 64), (128, 128, 64), (128, 64, 256), (256, 128, 512)]:
        a = torch.randn(M, K).to(torch.bfloat16)
        w = (torch.randn(N, K) / K ** 0.5).to(torch.bfloat16)
        ref = a.float() @ w.float().t()
        got = ops.linear_bf16(a.to(dev), w.to(dev)).cpu()
        rel = report(f'linear M={M} N={N} K={K}', got, ref)
        if rel > 1e-3 and K == 64:
            for j in range(4):                       # isolate each 16-wide K slice of the stage
                a2 = torch.zeros_like(a)
                a2[:, 16 * j:16 * j + 16] = a[:, 16 * j:16 * j + 16]
                report(f'   only K slice {j}', ops.linear_bf16(a2.to(dev), w.to(dev)).cpu(), a2.float() @ w.float().t())
            a3 = torch.zeros_like(a)
            a3[:, 0] = 1.0                           # D[m, n] = W[n, 0]
            got3 = ops.linear_bf16(a3.to(dev), w.to(dev)).cpu()
            print('   a[:,0]=1 -> got[0,:8]', got3[0, :8].tolist(), 'w[:8,0]', w[:8, 0].float().tolist())
    torch.cuda.synchronize()
    print('probe done')


if __name__ == '__main__':
    main()
