"""Per-kernel launch counts and time shares from an ncu launch list (--metrics gpu__time_duration.sum --csv).

    python tools/launch_shares.py gpurun_out/r02s_launches.csv > profiles/r02s_launch_shares.txt
"""
import collections
import csv
import sys


def main():
    rows = list(csv.reader(line for line in open(sys.argv[1]) if line.startswith('"')))
    hdr = rows[0]
    i_name, i_val, i_unit = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    scale = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3, "nsecond": 1e-6, "usecond": 1e-3, "msecond": 1.0, "second": 1e3}
    tot, cnt = collections.Counter(), collections.Counter()
    for r in rows[1:]:
        if len(r) != len(hdr):
            continue
        ms = float(r[i_val].replace(",", "")) * scale.get(r[i_unit], 1e-6)
        tot[r[i_name]] += ms
        cnt[r[i_name]] += 1
    total = sum(tot.values())
    print("# per-kernel launch counts and time shares of `python bench.py --steps 1 --warmup 3 --no-graph ...` under ncu "
          "(gpu__time_duration.sum,")
    print("# --clock-control none; per-launch times are cold-cache and serialised: shares, not absolutes).  4 infer calls = "
          "3 warm-ups + 1 timed.")
    print(f"# total {total:.1f} ms over {sum(cnt.values())} launches")
    for name, ms in tot.most_common():
        print(f"{100 * ms / total:6.2f} %  {ms:10.2f} ms  {cnt[name]:5d} launches  avg {ms / cnt[name]:8.4f} ms  {name[:110]}")


if __name__ == "__main__":
    main()
