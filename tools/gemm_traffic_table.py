"""profiles/r02_gemm_traffic_table.json from an `ncu --set full` capture of tools/ncu_targets.py (run where ncu is installed):

    python tools/gemm_traffic_table.py gpurun_out/r02_kernels_hd80.ncu-rep <call id> > profiles/r02_gemm_traffic_table.json

One row per GEMM launch of the capture (C2 block shapes, T = 8192 tokens): measured DRAM bytes (dram__bytes_read.sum +
dram__bytes_write.sum), duration, tensor-pipe activity, next to the ALGORITHMIC bytes of that launch (every operand read once,
every output written once, accumulated outputs also read once).  bench.py copies the dominant launch's two numbers into
`roofline.traffic` / `roofline.traffic_algorithmic`."""
import csv
import io
import json
import subprocess
import sys

T, H, F = 8192, 2560, 10240
# launches of tools/ncu_targets.py in order: forward c_fc, dgrad of c_fc, the four weight gradients of a block (accumulating)
SHAPES = [
    ("fwd c_fc  [T,2F] = x[T,H] W[2F,H]^T", (T * H + 2 * F * H) * 2 + T * 2 * F * 2),
    ("dgrad c_fc  [T,H] = dy[T,2F] W[2F,H]", (T * 2 * F + 2 * F * H) * 2 + T * H * 2),
    ("wgrad x4 (c_proj_mlp, c_fc, c_proj_attn, c_attn), fp32 accumulate",
     sum(T * (m + n) * 2 + m * n * 8 for m, n in ((H, F), (2 * F, H), (H, H), (3 * H, H)))),
]


def main():
    rep, call = sys.argv[1], sys.argv[2]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr = rows[0]
    idx = {h: i for i, h in enumerate(hdr)}
    gemms = [r for r in rows[2:] if "gemm_bf16_kernel" in r[idx["Kernel Name"]]]

    def num(r, k):
        return float(r[idx[k]].replace(",", ""))

    def unit_scale(k):
        u = rows[1][idx[k]]
        return {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ms": 1.0, "us": 1e-3, "%": 1.0}.get(u, 1.0)

    table = []
    for r, (what, alg) in zip(gemms, SHAPES):
        rd = num(r, "dram__bytes_read.sum") * unit_scale("dram__bytes_read.sum")
        wr = num(r, "dram__bytes_write.sum") * unit_scale("dram__bytes_write.sum")
        table.append({
            "launch": what, "kernel": r[idx["Kernel Name"]].split("(")[0].replace("void <unnamed>::", ""),
            "duration_ms": num(r, "gpu__time_duration.sum") * unit_scale("gpu__time_duration.sum"),
            "dram_bytes": rd + wr, "dram_read": rd, "dram_write": wr, "algorithmic_bytes": alg, "ratio": (rd + wr) / alg,
            "tensor_pipe_pct": num(r, "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed"),
        })
    dom = max(table, key=lambda t: t["duration_ms"])
    print(json.dumps({
        "source": f"gpurun call {call}: ncu --set full --clock-control none of tools/ncu_targets.py ({rep.split('/')[-1]}); "
                  "cold-cache serialised launches, T = 8192 tokens (the bench step runs the same kernels at T = 24576)",
        "dominant_launch": dom["launch"], "dominant_launch_dram_bytes": dom["dram_bytes"],
        "dominant_launch_algorithmic_bytes": dom["algorithmic_bytes"], "launches": table}, indent=1))


if __name__ == "__main__":
    main()
