"""Copy the artefacts of tools/gpu_official.sh from gpurun_out/ into profiles/ under a round/version tag."""
import csv, io, json, os, shutil, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r01_v7"
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
for src, dst in (("bench.json", "bench.json"), ("bench_reference.json", "bench_reference.json"), ("launches.csv", "launches.csv"),
                 ("nproc.txt", "host_cpu.txt")):
    if os.path.exists(os.path.join(G, src)):
        shutil.copy(os.path.join(G, src), os.path.join(P, f"{tag}_{dst}"))
rep = os.path.join(G, "prof_walk64.ncu-rep")
if os.path.exists(rep):
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_top.py"), rep, "45"], capture_output=True, text=True).stdout
    open(os.path.join(P, f"{tag}_walk64_ncu_summary.txt"), "w").write(out)
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    r = list(csv.reader(io.StringIO(raw)))
    row = dict(zip(r[0], r[2])); units = dict(zip(r[0], r[1]))
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    rd = float(row["dram__bytes_read.sum"]) * scale[units["dram__bytes_read.sum"]]
    wr = float(row["dram__bytes_write.sum"]) * scale[units["dram__bytes_write.sum"]]
    tpath = os.path.join(P, "traffic.json")
    t = json.load(open(tpath)) if os.path.exists(tpath) else {}
    import re
    mode = 3 if re.search(r"k_walk_uniform<\(?(int)?\)?\s*3", row.get("Kernel Name", "")) else 1     # BRICK8 or PACKED8 walk
    t[f"k_walk_uniform<{mode}>@64x256^3"] = int(rd + wr)
    t["_source" if mode == 1 else "_source_brick8"] = (f"ncu --set full, one launch of the 64-instance walk ({tag}): dram__bytes_read.sum "
                                                        f"{rd / 1e9:.3f} GB + dram__bytes_write.sum {wr / 1e9:.3f} GB")
    sectors = float(row.get("lts__t_sectors_srcunit_tex_op_atom.sum") or 0) + float(row.get("lts__t_sectors_srcunit_tex_op_red.sum") or 0)
    if mode == 3 and sectors:
        t["atom_sectors_per_sample<3>@64x256^3"] = round(sectors / 209630528.0, 4)            # samples of the bench frame (bench.py prints them)
    json.dump(t, open(tpath, "w"), indent=1)
    print("traffic", rd, wr)
