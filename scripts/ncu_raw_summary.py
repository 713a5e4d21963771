#!/usr/bin/env python
"""Aggregate an `ncu --page raw --csv` dump (optionally .gz) of a `--set full` capture by (kernel, grid): launches, mean duration,
DRAM bytes, DRAM / tensor-pipe / issue-slot utilisation, registers, dynamic shared memory.  Markdown on stdout.
    python scripts/ncu_raw_summary.py gpurun_out/prof_raw.csv.gz [--gemm-traffic out.json]
--gemm-traffic: also write the DRAM traffic of the longest `gemm_tcgen05_kernel<256, 4, 0, 0>` launch (the 8000 x 768 x 3072 product, the
shape with the largest share of the step) with the hash of the GEMM sources, for bench.py's roofline.traffic."""
import collections, csv, gzip, hashlib, io, json, os, re, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
COLS = {"us": "gpu__time_duration.sum", "rd": "dram__bytes_read.sum", "wr": "dram__bytes_write.sum",
        "dram_rd_pct": "dram__bytes_read.sum.pct_of_peak_sustained_elapsed", "dram_wr_pct": "dram__bytes_write.sum.pct_of_peak_sustained_elapsed",
        "tensor": "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "issue": "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "warps": "sm__warps_active.avg.pct_of_peak_sustained_active", "regs": "launch__registers_per_thread",
        "smem": "launch__shared_mem_per_block_dynamic", "inst": "smsp__inst_executed.sum"}
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}


def gemm_src_sha16() -> str:
    h = hashlib.sha256()
    for name in ("gemm_tcgen05.cu", "ptx_sm100.cuh", "common.cuh", "common.cu"):
        with open(os.path.join(ROOT, "jiao-liao_speech_recognition_b200", "csrc", name), "rb") as f:
            h.update(f.read())
    return h.hexdigest()[:16]


def short(n: str) -> str:
    return re.sub(r"\(.*", "", n).replace("void ", "").replace("jl::", "")[:70]


def main():
    path = sys.argv[1]
    f = io.TextIOWrapper(gzip.open(path)) if path.endswith(".gz") else open(path)
    rows = list(csv.reader(f))
    hdr, units, data = rows[0], rows[1], rows[2:]
    ix = {k: hdr.index(v) for k, v in COLS.items() if v in hdr}
    kn, gs = hdr.index("Kernel Name"), hdr.index("Grid Size")

    def val(r, k):
        if k not in ix or r[ix[k]] in ("", "n/a"):
            return 0.0
        try:
            v = float(r[ix[k]].replace(",", ""))
        except ValueError:
            return 0.0
        return v * UNIT.get(units[ix[k]].split("/")[0], 1.0)

    agg = collections.OrderedDict()
    for r in data:
        d = agg.setdefault((short(r[kn]), r[gs]), collections.defaultdict(float))
        d["n"] += 1
        for k in COLS:
            d[k] += val(r, k)
    print("| kernel | grid | launches | us | DRAM read MB | DRAM written MB | DRAM % of peak | tensor pipe % | issue slots % | warps active % | regs | dyn smem KB | warp instr (M) |")
    print("|---|---|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|")
    for (name, grid), d in agg.items():
        n = d["n"]
        print(f"| `{name}` | {grid} | {int(n)} | {d['us'] / n:.1f} | {d['rd'] / n / 1e6:.1f} | {d['wr'] / n / 1e6:.1f} | {(d['dram_rd_pct'] + d['dram_wr_pct']) / n:.1f} | {d['tensor'] / n:.1f} | "
              f"{d['issue'] / n:.1f} | {d['warps'] / n:.1f} | {int(d['regs'] / n)} | {d['smem'] / n / 1e3:.0f} | {d['inst'] / n / 1e6:.2f} |")
    if "--gemm-traffic" in sys.argv:
        out = sys.argv[sys.argv.index("--gemm-traffic") + 1]
        cand = [r for r in data if "gemm_tcgen05_kernel<256, 4, 0, 0>" in r[kn]]
        if cand:
            r = max(cand, key=lambda r: val(r, "rd"))          # the product with the 49.2 MB A operand
            tj = {"shape_mnk": [8000, 768, 3072], "dram_bytes": int(val(r, "rd") + val(r, "wr")), "algorithmic_bytes": 66158592,
                  "kernel": "gemm_tcgen05_kernel<256, 4, 0, 0>", "duration_us_ncu": val(r, "us"), "tensor_pipe_pct": val(r, "tensor"),
                  "gemm_src_sha16": gemm_src_sha16(),
                  "source": f"{os.path.basename(path)} (ncu --set full --clock-control none, the launch of gemm_tcgen05_kernel<256, 4, 0, 0> with the largest DRAM read in one warm "
                            "fine-tune step = the FFN product [8000 x 768, K = 3072]; A 49.2 MB + B 4.7 MB + residual / output 12.3 MB algorithmic)"}
            with open(out, "w") as g:
                json.dump(tj, g, indent=1)


if __name__ == "__main__":
    main()
