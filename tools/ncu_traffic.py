"""Write profiles/ncu_traffic.json from a full ncu capture of the dominant evaluation launch: DRAM bytes per evaluated point, keyed by the
hash of the kernel source, so that bench.py reports `roofline.traffic` only while the capture still describes the code.
   python tools/ncu_traffic.py gpurun_out/prof_r2_eval_pde.ncu-rep 3801600"""
import csv, hashlib, json, os, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rep, points = sys.argv[1], int(sys.argv[2])
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units, vals = rows[0], rows[1], rows[2]


def metric(name):
    i = hdr.index(name)
    v = float(vals[i].replace(",", ""))
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[units[i]]


rd, wr = metric("dram__bytes_read.sum"), metric("dram__bytes_write.sum")
src = "scasml_gp_b200/csrc/gp_eval_tc.cu"
rec = {"capture": os.path.basename(rep) + " (ncu --set full, level-1 PDE-class launch of a C3 step)", "points": points,
       "dram_bytes_read": rd, "dram_bytes_write": wr, "dram_bytes_per_point": round((rd + wr) / points, 1),
       "kernel_source": src, "kernel_source_sha16": hashlib.sha256(open(os.path.join(ROOT, src), "rb").read()).hexdigest()[:16]}
json.dump(rec, open(os.path.join(ROOT, "profiles", "ncu_traffic.json"), "w"), indent=1)
print(rec)
