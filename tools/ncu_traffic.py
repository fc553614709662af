"""profiles/rNN_traffic.json for bench.py's roofline.traffic: dram__bytes_read.sum + dram__bytes_write.sum of ONE launch of the
dominant GEMM, read from an `ncu --set full` capture.  usage: ncu_traffic.py capture.ncu-rep out.json "note" """
import csv
import io
import json
import subprocess
import sys


def main(rep, out, note=""):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, val = rows[0], rows[1], rows[-1]
    ix = {h: i for i, h in enumerate(hdr)}

    def b(name):
        v = float(val[ix[name]].replace(",", ""))
        u = units[ix[name]].lower()
        return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}[u]
    j = {"kernel": val[ix["Kernel Name"]], "dram_bytes_per_launch": int(b("dram__bytes_read.sum") + b("dram__bytes_write.sum")),
         "dram_bytes_read": int(b("dram__bytes_read.sum")), "dram_bytes_write": int(b("dram__bytes_write.sum")),
         "duration_us": float(val[ix["gpu__time_duration.sum"]]), "capture": rep, "note": note}
    json.dump(j, open(out, "w"), indent=1)
    print(json.dumps(j))


if __name__ == "__main__":
    main(*sys.argv[1:4])
