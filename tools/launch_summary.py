"""Aggregates an `ncu --metrics gpu__time_duration.sum --csv` launch list of `bench.py --steps 1 --warmup 1 ...` into
profiles/launches_*_summary.csv: the launches of the LAST ModelInference.infer in the log (sample_query_kernel .. occlusion_kernel),
per kernel: launches, total microseconds, share of the serialised step."""
import collections
import csv
import re
import sys


def main():
    src, dst = sys.argv[1], sys.argv[2]
    rows = [r for r in csv.reader(l for l in open(src) if l.startswith('"'))]
    h = rows[0]
    ki, vi = h.index("Kernel Name"), h.index("Metric Value")
    seq = [(re.sub(r"^(void )?(dtk::)?", "", r[ki]).split("(")[0], float(r[vi].replace(",", "")) / 1000.0) for r in rows[1:]]
    starts = [i for i, (k, _) in enumerate(seq) if k.startswith("sample_query_kernel")]
    ends = [i for i, (k, _) in enumerate(seq) if k.startswith("occlusion_kernel")]
    b = ends[-1]
    a = max(i for i in starts if i < b)      # the last COMPLETE step (the capture may stop inside a later one)
    agg = collections.OrderedDict()
    for k, us in seq[a:b + 1]:
        agg.setdefault(k, [0, 0.0])
        agg[k][0] += 1; agg[k][1] += us
    tot = sum(v[1] for v in agg.values())
    with open(dst, "w") as f:
        f.write("# ncu launch list (--metrics gpu__time_duration.sum --clock-control none) of: python bench.py --steps 1 --warmup 1 "
                "--stages 0 --cpu-baseline 0 --multi 0 --torch-cuda-baseline 0 --second-head 0 --stream-probe 0\n")
        f.write(f"# one ModelInference.infer (config 2: T=50, 256 queries, C=1024) = the launches from sample_query_kernel to "
                f"occlusion_kernel of the last step in the log: {b - a + 1} launches, {tot / 1000:.2f} ms serialised.\n")
        f.write("# per-launch times under ncu are cold-cache and serialised: compare the SHARES with the kernels block of "
                "bench_r2_final.json, not the absolutes.\n")
        f.write("kernel,launches,us_total,share\n")
        for k, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f'"{k}",{n},{us:.1f},{us / tot:.4f}\n')


if __name__ == "__main__":
    main()
