"""A/B timing of the WN kernels on ONE GPU in ONE process (box-to-box clock variance is +-5 %, so variants
must be compared inside the same run).  Each variant is looped for ~`--seconds` back to back (steady state
under the power cap), timed with CUDA events; board power and SM clock are sampled through NVML meanwhile.

    python tools/bench_kernels.py [--batch 64] [--frames 860] [--seconds 1.5] [--out gpurun_out/kernels.json]
"""
import argparse
import json
import os
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from text2speech_b200 import _lib, synthetic as syn          # noqa: E402
from text2speech_b200.packing import PackedWaveGlow          # noqa: E402

DEV = torch.device("cuda:0")


class Nvml:
    def __init__(self):
        self.rows, self.stop, self.h = [], False, None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(torch.cuda.current_device())
        except Exception:                                     # noqa: BLE001
            self.nv = None

    def __enter__(self):
        self.rows, self.stop = [], False
        if self.nv:
            self.t = threading.Thread(target=self._run, daemon=True)
            self.t.start()
        return self

    def _run(self):
        while not self.stop:
            try:
                self.rows.append((self.nv.nvmlDeviceGetPowerUsage(self.h) / 1000.0,
                                  self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM)))
            except Exception:                                 # noqa: BLE001
                pass
            time.sleep(0.02)

    def __exit__(self, *exc):
        self.stop = True
        if self.nv:
            self.t.join()

    def summary(self):
        rows = self.rows[len(self.rows) // 3:]                # steady state: drop the first third
        if not rows:
            return {"power_w": None, "sm_mhz": None}
        return {"power_w": round(sum(r[0] for r in rows) / len(rows), 1),
                "sm_mhz": round(sum(r[1] for r in rows) / len(rows), 1)}


def loop(fn, seconds):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        fn()
    e1.record()
    e1.synchronize()
    n = max(5, int(seconds * 1e3 / (e0.elapsed_time(e1) / 3)))
    with Nvml() as nv:
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        e1.synchronize()
    return e0.elapsed_time(e1) / n, n, nv.summary()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--frames", type=int, default=860)
    ap.add_argument("--seconds", type=float, default=1.5)
    ap.add_argument("--only", default="")
    ap.add_argument("--hints", default="", help="comma list of gate_l2_hint values to A/B (default: 0,1,2,3,0 at d = 1 and 128)")
    ap.add_argument("--hint-dilation", type=int, default=8)
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "kernels.json"))
    args = ap.parse_args()
    b, t = args.batch, args.frames * 32
    cfg = syn.load_config()
    pk = PackedWaveGlow(syn.synthetic_state_dict(cfg, seed=1234, end_std=0.01), 12, 8, 512, 8, "bf16", DEV)
    fl = pk.flows[3]
    s = _lib.stream_ptr()
    bf = torch.bfloat16
    h0 = torch.randn((b, t, 512), device=DEV).to(bf)
    h1 = torch.empty_like(h0)
    cond = torch.randn((b, t, 640), device=DEV).to(bf)
    acts_all = (torch.rand((8, b, t, 512), device=DEV) * 2 - 1).to(bf)
    x = torch.randn((b, t, 8), device=DEV)
    steps = b * t
    gate_flop, res_flop, skip_flop = 2 * 2176 * 1024 * steps, 2 * 512 * 512 * steps, 2 * 4096 * 512 * steps
    variants = []
    for name in ("wgb_tc_wn_gate", "wgb_tc2_wn_gate"):
        for d in (1, 128):
            variants.append((f"{name} d={d}", gate_flop, 3328 * steps,
                             lambda name=name, d=d: _lib.call(name, h0, cond, fl["w_gate"][2], fl["b_gate"][2], acts_all[2], b, t, d, s)))
    for name in ("wgb_tc_wn_res", "wgb_tc2_wn_res"):
        variants.append((name, res_flop, 3072 * steps,
                         lambda name=name: _lib.call(name, acts_all[2], fl["w_res"][2], fl["b_res"][2], h0, h1, b, t,
                                                     *((t, None, None, 0) if name == "wgb_tc2_wn_res" else ()), s)))
    def gate_hinted(hint, d):
        def run():
            _lib.call("wgb_set_tuning", "gate_l2_hint", hint)
            _lib.call("wgb_tc2_wn_gate", h0, cond, fl["w_gate"][2], fl["b_gate"][2], acts_all[2], b, t, d, s)
            _lib.call("wgb_set_tuning", "gate_l2_hint", 1)
        return run
    for hint in (0, 1, 0, 1):
        variants.append((f"wgb_tc2_wn_gate d=8 l2_hint={hint}", gate_flop, 3328 * steps, gate_hinted(hint, 8)))

    def res_hinted(hint):
        def run():
            _lib.call("wgb_set_tuning", "res_l2_hint", hint)
            _lib.call("wgb_tc2_wn_res", acts_all[2], fl["w_res"][2], fl["b_res"][2], h0, h1, b, t, t, None, None, 0, s)
            _lib.call("wgb_set_tuning", "res_l2_hint", 0)
        return run
    for hint in (0, 1, 2, 3, 0):
        variants.append((f"wgb_tc2_wn_res l2_hint={hint}", res_flop, 3072 * steps, res_hinted(hint)))
    skip_row = torch.zeros((b * t, 8), device=DEV)
    variants.append(("wgb_tc2_wn_res + skip pass", res_flop, 3136 * steps,
                     lambda: _lib.call("wgb_tc2_wn_res", acts_all[2], fl["w_res"][2], fl["b_res"][2], h0, h1, b, t, t,
                                       fl["w_skip16_layers"][2], skip_row, 0, s)))
    skip_args = (acts_all, 8, fl["w_skip"], fl["w_end_t"], fl["b_end"], x, fl["w_mix_inv"], None, b, t, fl["n_half"], 0)
    nf = pk.flows[2]
    variants.append(("wgb_tc_wn_skip_end", skip_flop, (8 * 1024 + 64) * steps,
                     lambda: _lib.call("wgb_tc_wn_skip_end", *skip_args, s)))
    variants.append(("wgb_tc2_wn_skip_end", skip_flop, (8 * 1024 + 64) * steps,
                     lambda: _lib.call("wgb_tc2_wn_skip_end", *skip_args, None, None, 0, None, 0, s)))
    variants.append(("wgb_tc2_wn_skip_end + next start", skip_flop, (9 * 1024 + 64) * steps,
                     lambda: _lib.call("wgb_tc2_wn_skip_end", *skip_args, nf["w_start"], nf["b_start"], nf["n_half"], h1, t, s)))
    s16 = (acts_all, 8, fl["w_skip16"], fl["b_end"], x, fl["w_mix_inv"], None, b, t, fl["n_half"], 0)
    variants.append(("wgb_tc_wn_skip16_end", skip_flop, (8 * 1024 + 64) * steps,
                     lambda: _lib.call("wgb_tc_wn_skip16_end", *s16, None, None, 0, None, 0, None, None, s)))
    variants.append(("wgb_tc_wn_skip16_end + next start", skip_flop, (9 * 1024 + 64) * steps,
                     lambda: _lib.call("wgb_tc_wn_skip16_end", *s16, nf["w_start"], nf["b_start"], nf["n_half"], h1, t, None, None, s)))
    s16l = (acts_all[7], 1, fl["w_skip16_layers"][7], fl["b_end"], x, fl["w_mix_inv"], None, b, t, fl["n_half"], 0)
    variants.append(("wgb_tc_wn_skip16_end last layer + acc + next start", 0, (1024 + 96 + 1024) * steps,
                     lambda: _lib.call("wgb_tc_wn_skip16_end", *s16l, nf["w_start"], nf["b_start"], nf["n_half"], h1, t,
                                       skip_row, None, s)))
    if pk.has_mel:
        stack = torch.randn((b, args.frames, 320), device=DEV).to(bf)
        for d in (1, 128):
            variants.append((f"wgb_tc2_wn_gate_mel d={d}", gate_flop, (2048 + 20) * steps,
                             lambda d=d: _lib.call("wgb_tc2_wn_gate_mel", h0, stack, fl["w_gate"][2], fl["w_mel"][2],
                                                   fl["b_mel"][2], acts_all[2], b, t, args.frames, d, None, None, 0, s)))
        fp = args.frames + 4
        h_pad = torch.zeros((b, 32 * fp, 512), device=DEV, dtype=bf)
        h_pad[:, :t] = h0
        stack_pad = torch.randn((b, fp, 320), device=DEV).to(bf)
        variants.append(("wgb_tc2_wn_gate_mel padded d=128", gate_flop, (2048 + 20) * steps,
                         lambda: _lib.call("wgb_tc2_wn_gate_mel", h_pad, stack_pad, fl["w_gate"][2], fl["w_mel"][2],
                                           fl["b_mel"][2], acts_all[2], b, t, fp, 128, None, None, 0, s)))

        # L2 eviction-priority hints on the gate kernel's TMA loads (wgb_set_tuning "gate_l2_hint"): 1 = weights evict_last,
        # 2 = h taps evict_first, 3 = both; d = 1 and d = 128
        def hinted(hint, d):
            def run():
                _lib.call("wgb_set_tuning", "gate_l2_hint", hint)
                _lib.call("wgb_tc2_wn_gate_mel", h_pad, stack_pad, fl["w_gate"][2], fl["w_mel"][2], fl["b_mel"][2], acts_all[2],
                          b, t, fp, d, None, None, 0, s)
                _lib.call("wgb_set_tuning", "gate_l2_hint", 1)
            return run
        for d in (1, 128) if not args.hints else (args.hint_dilation,):
            for hint in (0, 1, 2, 3, 0) if not args.hints else [int(v) for v in args.hints.split(",")]:
                variants.append((f"wgb_tc2_wn_gate_mel padded d={d} l2_hint={hint}", gate_flop, (2048 + 20) * steps, hinted(hint, d)))
        skip_acc = torch.zeros((4, b * t, 8), device=DEV)
        variants.append(("wgb_tc2_wn_gate_mel padded + skip acc d=128", gate_flop, (2048 + 20 + 256) * steps,
                         lambda: _lib.call("wgb_tc2_wn_gate_mel", h_pad, stack_pad, fl["w_gate"][2], fl["w_mel"][2],
                                           fl["b_mel"][2], acts_all[2], b, t, fp, 128, fl["w_comp"][2], skip_acc, 0, s)))
        variants.append(("wgb_end_from_acc + next start", 0, (160 + 32 + 1024) * steps,
                         lambda: _lib.call("wgb_end_from_acc", skip_acc, fl["b_end"], x, fl["w_mix_inv"], None, b, t,
                                           fl["n_half"], 0, nf["w_start"], nf["b_start"], nf["n_half"], h_pad, 32 * fp, s)))
    variants.append(("wgb_wn_start", 0, (32 + 1024) * steps,
                     lambda: _lib.call("wgb_wn_start", x, fl["w_start"], fl["b_start"], h0, 1, b * t, 512, fl["n_half"], s)))
    out = []
    for name, flop, bytes_, fn in variants:
        if args.only and args.only not in name:
            continue
        ms, n, nv = loop(fn, args.seconds)
        rec = {"kernel": name, "ms": round(ms, 4), "iters": n, "tflops": round(flop / ms / 1e9, 1) if flop else None,
               "algorithmic_gb_s": round(bytes_ / ms / 1e6, 1), **nv}
        if nv["power_w"]:
            rec["joule_per_launch"] = round(nv["power_w"] * ms * 1e-3, 3)
            if flop and nv["sm_mhz"]:
                rec["frac_of_clock_peak"] = round(flop / (ms * 1e-3) / (148 * 8192 * nv["sm_mhz"] * 1e6), 3)
        out.append(rec)
        print(json.dumps(rec), flush=True)
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    with open(args.out, "w") as f:
        json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
