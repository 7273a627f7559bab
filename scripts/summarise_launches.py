"""Summarise an ncu --csv launch list (gpu__time_duration + dram bytes + tensor-pipe activity) per kernel and per launch.
usage: python scripts/summarise_launches.py gpurun_out/r01_launches_v3.csv STEPS > profiles/..."""
import collections
import csv
import sys


def load(path):
    with open(path) as f:
        lines = [l for l in f if l.startswith('"')]
    per = collections.OrderedDict()
    for d in csv.DictReader(lines):
        e = per.setdefault(d["ID"], {"name": d["Kernel Name"], "grid": d["Grid Size"], "block": d["Block Size"]})
        e[d["Metric Name"]] = float(d["Metric Value"].replace(",", ""))
    return list(per.values())


def short(name):
    name = name.replace("void ", "")
    return name.split("(")[0][:56]


def main():
    path, steps = sys.argv[1], int(sys.argv[2])
    L = load(path)
    T = "gpu__time_duration.sum"
    tot = sum(e[T] for e in L)
    print(f"# {path}: {len(L)} launches over {steps} bench steps ({len(L) // steps} per step), ncu --clock-control none, per-launch metrics")
    print("# per-launch times are cold-cache and serialised: compare SHARES with bench.py's live CUDA-event times, not absolutes")
    print(f"# total {tot / 1e6:.3f} ms")
    agg = collections.OrderedDict()
    for e in L:
        a = agg.setdefault(short(e["name"]), [0, 0.0, 0.0, 0.0, 0.0])
        a[0] += 1
        a[1] += e[T]
        a[2] += e.get("dram__bytes_read.sum", 0)
        a[3] += e.get("dram__bytes_write.sum", 0)
        a[4] += e.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", 0) * e[T]
    print(f"{'kernel':58s} {'n/step':>6s} {'us/step':>9s} {'share':>6s} {'dramR MB/step':>13s} {'dramW MB/step':>13s} {'tensor-pipe %':>13s}")
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{k:58s} {a[0] // steps:6d} {a[1] / steps / 1e3:9.1f} {100 * a[1] / tot:5.1f}% {a[2] / steps / 1e6:13.1f} {a[3] / steps / 1e6:13.1f} {a[4] / a[1]:13.1f}")
    print()
    print("# launches of the LAST step, in order")
    n = len(L) // steps
    print(f"{'#':>3s} {'kernel':58s} {'grid':>12s} {'us':>8s} {'dramR MB':>9s} {'dramW MB':>9s} {'GB/s':>7s} {'tensor %':>8s} {'SM GHz':>6s}")
    for i, e in enumerate(L[-n:]):
        r, w = e.get("dram__bytes_read.sum", 0), e.get("dram__bytes_write.sum", 0)
        print(f"{i:3d} {short(e['name']):58s} {e['grid']:>12s} {e[T] / 1e3:8.1f} {r / 1e6:9.1f} {w / 1e6:9.1f} {(r + w) / e[T]:7.0f} "
              f"{e.get('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed', 0):8.1f} {e.get('sm__cycles_elapsed.avg.per_second', 0) / 1e9:6.2f}")
    conv = [e for e in L[-n:] if "conv_v3" in e["name"] or "conv_igemm" in e["name"] or "conv_strip" in e["name"]]
    cr = sum(e.get("dram__bytes_read.sum", 0) for e in conv)
    cw = sum(e.get("dram__bytes_write.sum", 0) for e in conv)
    print(f"\n# conv launches of one step: {len(conv)}, time {sum(e[T] for e in conv) / 1e3:.1f} us, DRAM read {cr / 1e6:.1f} MB + write {cw / 1e6:.1f} MB = {(cr + cw)} bytes")


if __name__ == "__main__":
    main()
