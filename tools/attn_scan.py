import sys, os
sys.path.insert(0, "tests"); sys.path.insert(0, ".")
import gpu_diag as D
for (B,H,n) in [(32,16,1875),(8,16,7500),(64,16,940),(32,16,1024),(32,16,2048)]:
    D.bench_attn(B,H,n,iters=5)
