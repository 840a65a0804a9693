"""Developer tool: hottest source lines (warp-stall samples) of an ncu report captured with --import-source on.
    python -m tools.ncu_hot_lines gpurun_out/prof.ncu-rep [top]
"""
import csv
import io
import subprocess
import sys


def main(path, top=40):
    out = subprocess.run(["ncu", "-i", path, "--page", "source", "--print-source", "cuda,sass", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    cur_file, items, total = None, [], 0
    hdr = None
    for r in rows:
        if len(r) == 2 and r[0] == "File Path":
            cur_file = r[1].split("/")[-1]
        elif len(r) > 10 and r[0] == "Line No":
            hdr = r
        elif len(r) > 10 and r[0] not in ("", "Line No"):
            try:
                samples = int(r[4] or 0)
            except ValueError:
                continue
            inst = int(r[hdr.index("Instructions Executed")] or 0) if hdr and "Instructions Executed" in hdr else 0
            items.append((samples, inst, cur_file, r[0], r[1].strip()[:110]))
            total += samples
    print(f"total samples {total}")
    for s, inst, f, ln, src in sorted(items, reverse=True)[:top]:
        print(f"{s:7d} {100.0 * s / max(total, 1):5.1f}%  inst {inst:9d}  {f}:{ln}  {src}")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40)
