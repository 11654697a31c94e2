#!/usr/bin/env python
"""Turn the raw captures under gpurun_out/ (scratch) into the tracked evidence files under profiles/ (dev tool).

    python tools/make_profiles.py ncu   NAME.ncu-rep [...]  -> text summary of the named `ncu --set full` captures on stdout
    python tools/make_profiles.py sass                      -> SASS mnemonic counts + excerpts of the tensor-core kernels
    python tools/make_profiles.py hbm   NAME.ncu-rep PLAIN.log -> table of the HBM-bound kernels (tools/gpu_ew_once.py capture)
"""
import csv
import io
import re
import subprocess
import sys

METRICS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
           "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
           "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
           "launch__registers_per_thread", "launch__grid_size", "launch__cluster_size", "launch__shared_mem_per_block_dynamic",
           "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
           "smsp__cycles_active.avg", "smsp__inst_executed.sum", "lts__t_sector_hit_rate.pct"]


def ncu_summary(path, flops=None):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        print("== " + r[hdr.index("Kernel Name")])
        vals = {}
        for m in METRICS:
            if m in hdr:
                i = hdr.index(m)
                vals[m] = r[i]
                print(f"   {m:75s} {r[i]} {units[i]}")
        if flops:
            t = float(vals["gpu__time_duration.sum"].replace(",", ""))
            u = units[hdr.index("gpu__time_duration.sum")]
            ms = t / 1e3 if u in ("us", "usecond") else (t / 1e6 if u in ("ns", "nsecond") else t)
            print(f"   -> {flops / ms / 1e9:.0f} TFLOP/s under ncu (cold cache, serialised)")


def sass():
    lib = "vae_gan_mark_b200/libvaegan_b200.so"
    txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
    pat = re.compile(r"\b(UTCHMMA(?:\.2CTA)?|UTMALDG\.[0-9]D(?:\.2CTA)?|UTCBAR(?:\.2CTA)?(?:\.MULTICAST)?|LDTM[.A-Za-z0-9]*|UTCATOMSWS[.A-Z_0-9]*|ELECT|R2UR(?:\.BROADCAST)?|BRA\.U\.ANY)\b")
    counts = {}
    for m in pat.finditer(txt):
        counts[m.group(1)] = counts.get(m.group(1), 0) + 1
    print("Mnemonic counts over the whole library (`cuobjdump -sass " + lib + "`, nvcc 12.9, sm_100a only):")
    for k in sorted(counts):
        print(f"  {k:36s} {counts[k]}")
    print("  (BRA.U.ANY = the ELECT / R2UR.BROADCAST waterfall loops the compiler wraps around a tcgen05 / TMA instruction issued from a\n"
          "   per-lane region: 0 since the issue loops are warp-uniform; before, every UTCHMMA / UTMALDG sat in one)\n")
    cur, body = None, {}
    for line in txt.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            continue
        if cur and ("conv_fprop_kernel" in cur or "conv_wgrad_kernel" in cur) and re.search(r"/\*[0-9a-f]{4,}\*/", line):
            body.setdefault(cur, []).append(line.rstrip())
    for fn, lines in body.items():
        print("## " + fn)
        idx = [i for i, ln in enumerate(lines) if "UTCHMMA.2CTA" in ln] or [i for i, ln in enumerate(lines) if "UTCHMMA" in ln]
        # the K loop of the plain mode: the densest run of 4 MMAs, with the instructions between them
        best = None
        for a in range(len(idx) - 3):
            if best is None or idx[a + 3] - idx[a] < best[1] - best[0]:
                best = (idx[a], idx[a + 3])
        for ln in lines[max(0, best[0] - 14):best[1] + 4]:
            print(re.sub(r"\s+/\* 0x[0-9a-f]+ \*/\s*$", "", ln))
        print()


def hbm(path, plain_log):
    """One line per profiled launch of tools/gpu_ew_once.py: duration, DRAM bytes and achieved DRAM bandwidth under ncu, next
    to the CUDA-event numbers of the plain run of the same script."""
    import json
    try:
        peak = json.load(open("MEASURED_PEAKS.json"))["hbm_gbs"]
    except Exception:
        peak = 6650.0
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]

    def val(r, m, scale):
        i = hdr.index(m)
        return float(r[i].replace(",", "")) * scale[units[i]]
    t_s = {"ms": 1e-3, "us": 1e-6, "ns": 1e-9, "msecond": 1e-3, "usecond": 1e-6, "nsecond": 1e-9, "s": 1.0, "second": 1.0}
    b_s = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "Tbyte": 1e12}
    one = {u: 1.0 for u in set(units)}
    print(f"copy peak of the pool (MEASURED_PEAKS.json hbm_gbs) = {peak:.0f} GB/s; `dram %` is ncu's gpu__dram_throughput against the\n"
          f"hardware peak (a plain copy reaches ~80 % on that scale); DRAM GB/s = (dram__bytes_read + dram__bytes_write) / duration\n")
    print(f"{'kernel':44s} {'us':>8s} {'read MB':>9s} {'write MB':>9s} {'DRAM GB/s':>10s} {'of copy':>8s} {'dram %':>7s} {'L2 hit %':>8s} {'regs':>5s} {'grid':>6s}")
    for r in rows[2:]:
        name = re.sub(r"\(.*", "", r[hdr.index("Kernel Name")]).replace("void ", "").replace("__nv_bfloat16", "bf16")
        t = val(r, "gpu__time_duration.sum", t_s)
        rd, wr = val(r, "dram__bytes_read.sum", b_s), val(r, "dram__bytes_write.sum", b_s)
        gbs = (rd + wr) / t / 1e9
        print(f"{name:44s} {t * 1e6:8.1f} {rd / 1e6:9.1f} {wr / 1e6:9.1f} {gbs:10.0f} {gbs / peak:8.2f} "
              f"{val(r, 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', one):7.1f} {val(r, 'lts__t_sector_hit_rate.pct', one):8.1f} "
              f"{int(val(r, 'launch__registers_per_thread', one)):5d} {int(val(r, 'launch__grid_size', one)):6d}")
    print("\nCUDA events, plain run of the same script (5 back-to-back launches per kernel, warm; algorithmic bytes = every tensor once):")
    for ln in open(plain_log):
        if ln.startswith("{"):
            d = json.loads(ln)
            print(f"  {d['kernel']:38s} {str(d['shape']):22s} {d['algorithmic_MB']:8.1f} MB {d['ms'] * 1e3:8.1f} us {d['GB/s']:8.0f} GB/s  {d['frac_of_copy_peak']:.2f} of copy peak")


if __name__ == "__main__":
    if sys.argv[1] == "hbm":
        hbm(sys.argv[2], sys.argv[3])
    elif sys.argv[1] == "ncu":
        fl = 2.0 * 64 * 128 * 128 * 512 * 4608
        for pth in sys.argv[2:]:
            ncu_summary(pth, fl if "film4" in pth else None)
    else:
        sass()
