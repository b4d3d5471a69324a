"""profiles/<round>_ncu_traffic.json from an `ncu --set full` report of profiles/run_infer_once.py B 2: DRAM bytes, tensor
pipe and duration of the LAST launch of every inference kernel, keyed by the op name prefix bench.py uses.
    python profiles/ncu_traffic.py X.ncu-rep B > profiles/r2_ncu_traffic.json"""
import csv, io, json, subprocess, sys

rep, images = sys.argv[1], int(sys.argv[2])
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[0]
ix = {h: i for i, h in enumerate(hdr)}


def f(r, k):
    try:
        return float(r[ix[k]].replace(",", ""))
    except (KeyError, ValueError):
        return 0.0


def key_of(name, grid):
    g = grid.replace(" ", "")
    if "conv1" in name or "tcp1" in name:
        return f"tc_layer[kind0,{images}x512x768"
    if "deconv3x" in name:
        return f"tc_layer[kind4,{images}x128x192"
    if "quant_entropy" in name:
        return "quant_entropy_infer"
    if "igemm_tcp_kernel<1, 1, 1" in name or "igemm_tcp_kernel<1, true, 1" in name:
        return f"tc_layer[kind1,{images}x128x192"
    if "igemm_tcp_kernel<1" in name:
        return f"tc_layer[kind1,{images}x64x96"
    if "igemm_tcp_kernel<2" in name:
        return f"tc_layer[kind2,{images}x32x48|{images}x64x96"      # resolved by order below
    return None


out, seen2 = {}, 0
for r in rows[2:]:
    name, grid = r[ix["Kernel Name"]], r[ix["Grid Size"]]
    k = key_of(name, grid)
    if k is None:
        continue
    if "|" in k:                       # the two transposed 5x5 layers run in order deconv1, deconv2 in every forward
        k = (f"tc_layer[kind2,{images}x32x48", f"tc_layer[kind2,{images}x64x96")[seen2 % 2]
        seen2 += 1
    rd, wr = f(r, "dram__bytes_read.sum"), f(r, "dram__bytes_write.sum")
    unit_rd = rows[1][ix["dram__bytes_read.sum"]]
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit_rd, 1.0)
    unit_t = rows[1][ix["gpu__time_duration.sum"]]
    tscale = {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(unit_t, 1.0)
    out[k] = {"dram_mbytes_per_launch": (rd + wr) * scale / 1e6, "dram_read_mbytes": rd * scale / 1e6,
              "dram_write_mbytes": wr * scale / 1e6,
              "tensor_pipe_active_pct": f(r, "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"),
              "duration_us": f(r, "gpu__time_duration.sum") * tscale, "ncu_kernel": name[:60]}
json.dump({"source": "ncu --set full --clock-control none --import-source on; profiles/run_infer_once.py "
                     f"{images} 2 (the bench batch; last launch of each kernel); dram__bytes_read.sum + "
                     "dram__bytes_write.sum per launch", "images": images, "kernels": out}, sys.stdout, indent=1)
