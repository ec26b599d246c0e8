"""First-contact diagnostics for the tcgen05 pipeline (run on the GPU box; writes gpurun_out/diag.json).

Exact-arithmetic GEMMs (small integers, exactly representable in bf16, fp32 sums exact) so any
mismatch is a layout / descriptor / pipeline bug rather than rounding; on mismatch the error is
summarised per 32x32 block to show the pattern.
"""
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from text2speech_b200 import _lib  # noqa: E402

DEV = "cuda:0"
out = {"cases": []}


def case(batch, T, N, K, seed):
    g = torch.Generator().manual_seed(seed)
    a = torch.randint(-3, 4, (batch, T, K), generator=g).float()
    w = torch.randint(-3, 4, (N, K), generator=g).float()
    bias = torch.randint(-8, 9, (N,), generator=g).float()
    want = a.double() @ w.double().t() + bias.double()
    c = torch.full((batch, T, N), float("nan"), device=DEV)
    t0 = time.time()
    try:
        _lib.call("wgb_tc_gemm", a.to(DEV, torch.bfloat16), w.to(DEV, torch.bfloat16), bias.to(DEV), c, 0, batch, T, N, K,
                  _lib.stream_ptr())
        torch.cuda.synchronize()
    except Exception as e:  # noqa: BLE001
        out["cases"].append({"shape": [batch, T, N, K], "error": repr(e)})
        return False
    got = c.cpu().double()
    bad = ~torch.isclose(got, want, atol=1e-3, rtol=0)
    rec = {"shape": [batch, T, N, K], "ms": (time.time() - t0) * 1e3, "n_bad": int(bad.sum()), "n": bad.numel(),
           "nan": int(torch.isnan(got).sum())}
    if rec["n_bad"]:
        b0 = bad[0]
        tb, nb = (T + 31) // 32, N // 32
        blk = torch.zeros(tb, nb)
        for i in range(tb):
            for j in range(nb):
                blk[i, j] = b0[i * 32:(i + 1) * 32, j * 32:(j + 1) * 32].float().mean()
        rec["bad_frac_per_32x32_block_batch0"] = [[round(float(v), 2) for v in row] for row in blk[:8]]
        idx = bad.nonzero()[:8].tolist()
        rec["examples"] = [{"idx": i, "got": float(got[tuple(i)]), "want": float(want[tuple(i)])} for i in idx]
    out["cases"].append(rec)
    return rec["n_bad"] == 0


def main():
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    _lib.require_b200(torch.device(DEV))
    ok = True
    for shape in [(1, 128, 256, 64, 1), (1, 128, 256, 256, 2), (1, 128, 512, 128, 3), (1, 100, 256, 64, 4),
                  (2, 300, 512, 192, 5), (3, 1000, 1024, 2176, 6)]:
        ok = case(*shape) and ok
        with open(os.path.join(ROOT, "gpurun_out", "diag.json"), "w") as f:
            json.dump(out, f, indent=1)
    print(json.dumps(out)[:3000])
    print("DIAG", "OK" if ok else "FAILED")
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
