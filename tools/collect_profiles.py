"""Copy the measurements of the final round-2 GPU runs from gpurun_out/ (scratch) into profiles/ (tracked) and extract
profiles/r02_ncu_traffic.json (DRAM bytes and time per launch from the `ncu --set full` raw pages; bench.py reports `roofline.traffic` from it).
    python tools/collect_profiles.py"""
import csv
import json
import os
import shutil

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC, DST = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
COPY = {
    # final 1-GPU runs (tools/r02_gpu_j.sh: final code; tools/r02_gpu_i.sh: the captures that do not depend on the last change)
    "r02j_bench_n1.json": "r02_bench_n1.json", "r02i_bench_reference.json": "r02_bench_reference.json",
    "r02j_bench_n1_two_buffers.json": "r02_bench_n1_two_buffers.json", "r02j_bench_n1_save_every_step.json": "r02_bench_n1_save_every_step.json",
    "r02j_bench_81x161x81.json": "r02_bench_81x161x81.json", "r02j_bench_512.json": "r02_bench_512.json",
    "r02i_transient_81x161x81_nt200_1.json": "r02_transient_81x161x81_nt200_run1.json", "r02i_transient_81x161x81_nt200_2.json": "r02_transient_81x161x81_nt200_run2.json",
    "r02i_transient_81x161x81_nt200_budget8GB.json": "r02_transient_81x161x81_nt200_budget8GB.json",
    "r02j_launches_bench_default.csv": "r02_launches_bench_default.csv", "r02j_launches_bench_81x161x81.csv": "r02_launches_bench_81x161x81.csv",
    "r02j_tests.log": "r02_gpu_tests.log",
    # multi-GPU runs (tools/r02_gpu_f.sh: 2 GPUs, tools/r02_gpu_g.sh: 8 GPUs)
    "r02f_bench_n2.json": "r02_bench_n2.json", "r02f_bench_n2_reference.json": "r02_bench_n2_reference_arm_under_torchrun.json", "r02f_bench_n2_strong512.json": "r02_bench_n2_strong512.json",
    "r02f_tests.log": "r02_gpu_tests_2gpus.log",
    "r02g_bench_n8.json": "r02_bench_n8.json", "r02g_bench_n8_heatsink3d.json": "r02_bench_n8_config3_heatsink3d_81x161x81_pe222.json",
    "r02g_bench_n8_strong512.json": "r02_bench_n8_strong512.json", "r02g_transient_81x161x81_nt200_pe222.json": "r02_transient_81x161x81_nt200_pe222_8gpus.json",
    # experiments kept as evidence (profiles/r02_tuning.md)
    "r02c_bench_n1.json": "r02_exp_bench_n1_cp_async_pipeline.json", "r02c_bench_n1_occ5_nopipe.json": "r02_exp_bench_n1_occ5_96_registers.json",
    "r02c_ncu_full_fused_fwd_gather_raw.csv": "r02_exp_ncu_full_fused_pipe_fwd_gather_raw.csv",
    "r02c_ncu_full_fused_fwd_local_nopipe_raw.csv": "r02_exp_ncu_full_fused_fwd_local_no_l2_ahead_raw.csv",
    "r02h_bench_41x81x41.json": "r02_exp_bench_41x81x41_cooperative.json", "r02j_bench_41x81x41.json": "r02_bench_41x81x41.json",
    "r02k_tests_2gpus.log": "r02_gpu_tests_final_2gpus.log", "r02k_tests_8gpus.log": "r02_gpu_tests_final_8gpus.log",
    "r02h_bench_n1.json": "r02_exp_bench_n1_small_domains_cooperative.json", "r02h_bench_n1_nocoop.json": "r02_exp_bench_n1_small_domains_launches.json",
}
for k in ("fwd_gather", "fwd_local", "adj_gather", "adj_local"):
    COPY[f"r02j_ncu_full_fused_{k}_raw.csv"] = f"r02_ncu_full_fused_{k}_raw.csv"
for k in ("fwd_storing", "adj_storing", "ns"):
    COPY[f"r02i_ncu_full_fused_{k}_raw.csv"] = f"r02_ncu_full_fused_{k}_raw.csv"
for k in ("k_xclose", "k_shell", "k_tubes", "k_sensitivity", "k_filter", "k_residual"):
    COPY[f"r02i_ncu_full_{k}_raw.csv"] = f"r02_ncu_full_{k}_raw.csv"


def main():
    missing = []
    for a, b in COPY.items():
        p = os.path.join(SRC, a)
        if os.path.exists(p) and os.path.getsize(p) > 0:
            shutil.copyfile(p, os.path.join(DST, b))
        else:
            missing.append(a)
    # L2 prefetch distance sweep (tools/r02_gpu_e.sh) in one file
    sweep = {}
    for size in ("n1", "81x161x81"):
        for a in (0, 18, 37, 74, 111, 148, 185, 222, 296, 592, 1184):
            for pre in ("r02e", "r02d"):
                p = os.path.join(SRC, f"{pre}_bench_{size}_ahead{a}.json")
                if os.path.exists(p) and os.path.getsize(p) > 0:
                    d = json.load(open(p))
                    sweep.setdefault(size, {})[str(a)] = {"value": d["value"], "forward_mlups": d["sweeps"]["forward_mlups"], "adjoint_mlups": d["sweeps"]["adjoint_mlups"],
                                                         "roofline_frac": d["roofline"]["frac"], "roofline_adjoint_frac": d["roofline_adjoint"]["frac"]}
                    break
    json.dump({"what": "bench.py --steps 20 --warmup 5 with PANSLBM_L2_AHEAD = distance in CTAs (in place; 352^3 = n1, 81x161x81)", "sweep": sweep},
              open(os.path.join(DST, "r02_l2_ahead_sweep.json"), "w"), indent=1)
    # DRAM traffic per launch
    kernels = {}
    names = {"r02_ncu_full_fused_fwd_gather_raw.csv": "k_fused<3,7>/elided/gather", "r02_ncu_full_fused_fwd_local_raw.csv": "k_fused<3,7>/elided",
             "r02_ncu_full_fused_fwd_storing_raw.csv": "k_fused<3,7>", "r02_ncu_full_fused_adj_gather_raw.csv": "k_fused<3,11>/elided/gather",
             "r02_ncu_full_fused_adj_local_raw.csv": "k_fused<3,11>/elided", "r02_ncu_full_fused_adj_storing_raw.csv": "k_fused<3,11>",
             "r02_ncu_full_k_xclose_raw.csv": "k_xclose<3,1>", "r02_ncu_full_k_shell_raw.csv": "k_shell<3,7>", "r02_ncu_full_k_tubes_raw.csv": "k_tubes<3,7>",
             "r02_ncu_full_k_sensitivity_raw.csv": "k_sensitivity<3>", "r02_ncu_full_k_filter_raw.csv": "k_filter", "r02_ncu_full_k_residual_raw.csv": "k_residual_partial"}
    units = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "ms": 1e3, "us": 1.0, "ns": 1e-3, "s": 1e6, "msecond": 1e3, "usecond": 1.0}

    def read(path):
        rows = list(csv.reader(open(path)))
        if len(rows) < 3:
            return []
        hdr, un = rows[0], rows[1]
        out = []
        for r in rows[2:]:
            d = dict(zip(hdr, r))
            u = dict(zip(hdr, un))
            val = lambda k: float(d[k].replace(",", ""))*units.get(u[k], 1.0)
            out.append({"name": d.get("Kernel Name", ""), "grid_threads": int(float(d["launch__grid_size"]))*int(float(d["launch__block_size"])),
                        "dram_bytes_per_launch": val("dram__bytes_read.sum") + val("dram__bytes_write.sum"), "dram_bytes_read": val("dram__bytes_read.sum"),
                        "dram_bytes_write": val("dram__bytes_write.sum"), "time_us": val("gpu__time_duration.sum"), "registers": int(float(d["launch__registers_per_thread"]))})
        return out
    for f, key in names.items():
        p = os.path.join(DST, f)
        if os.path.exists(p):
            r = read(p)
            if r:
                kernels[key] = r[0]
    p = os.path.join(DST, "r02_ncu_full_fused_ns_raw.csv")
    if os.path.exists(p):
        for r in read(p):
            key = "k_fused<3,1>/elided" + ("/gather" if "<3, 1, 1>" in r["name"] or "ELi1EEE" in r["name"] else "")
            kernels.setdefault(key, r)
    json.dump({"source": "ncu --set full --clock-control none, one launch each, bench.py default sizes (352^3 heatsink sweep, 512^3 NS cavity), tools/r02_gpu_i.sh; "
                         "raw pages in profiles/r02_ncu_full_*_raw.csv.  '/elided' = the pass that stores on the closure planes only (local pass unless '/gather').",
               "kernels": kernels}, open(os.path.join(DST, "r02_ncu_traffic.json"), "w"), indent=1)
    print("copied", len(COPY) - len(missing), "missing", missing)
    print(json.dumps({k: (round(v["dram_bytes_per_launch"]/1e9, 3), round(v["time_us"], 1), v["registers"]) for k, v in kernels.items()}, indent=0))


if __name__ == "__main__":
    main()
