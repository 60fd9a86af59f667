"""Run on the GPU box (under gpurun): launch list of one factorization, then `ncu --set full`
of the longest k_gemm_batched launches (one contribution update, one trailing update).
Outputs land in gpurun_out/."""
import csv
import subprocess
import sys

CASE = sys.argv[1:] or ["lap27", "100"]
RUN = ["python", "scripts/run_case.py", *CASE, "--reps", "2", "--nosolve"]
out = "gpurun_out"
subprocess.run(["ncu", "--metrics", "gpu__time_duration.sum", "--clock-control", "none", "--csv",
                "--log-file", f"{out}/launches.csv", *RUN], check=False, stdout=subprocess.DEVNULL)
rows = [r for r in csv.reader(open(f"{out}/launches.csv")) if len(r) > 10 and r[0].isdigit()]
names = [r[4].split("(")[0] for r in rows]
times = [float(r[-1]) for r in rows]
half = len(rows) // 2                      # second factorization = second half of the launches
gemm = [(i, t) for i, (nm, t) in enumerate(zip(names, times)) if nm == "k_gemm_batched"]
gemm2 = [(i, t) for i, t in gemm if i >= half]
longest = max(gemm2, key=lambda x: x[1])
# ordinal among k_gemm_batched launches (for -k ... -s)
ordinal = [i for i, _ in gemm].index(longest[0])
print("launches", len(rows), "gemm launches", len(gemm), "longest gemm ordinal", ordinal, "ns", longest[1])
# the longest launch and the longest of the last 60 (root front trailing updates)
tail = max(gemm2[-60:], key=lambda x: x[1])
ord_tail = [i for i, _ in gemm].index(tail[0])
for tag, o in (("contrib", ordinal), ("update", ord_tail)):
    subprocess.run(["ncu", "--set", "full", "--clock-control", "none", "--import-source", "on", "-k",
                    "regex:k_gemm_batched", "-s", str(o), "-c", "1", "-f", "-o", f"{out}/prof_gemm_{tag}", *RUN],
                   check=False, stdout=subprocess.DEVNULL)
print("done")
