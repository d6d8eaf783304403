"""Per-source-line executed-instruction and stall-sample totals from an ncu report captured with
--import-source on.   python profiles/tools/source_hotspots.py report.ncu-rep [top_n]"""
import csv
import subprocess
import sys


def main():
    rep = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    cur = line = src = None
    tot, samp, text, calls, seen = {}, {}, {}, [], set()
    for r in rows:
        if len(r) >= 2 and r[0] == "File Path":
            cur = r[1].split("/")[-1]
            continue
        if len(r) < 8 or r[0] in ("Line No", "Function Name"):
            continue
        if r[0] != "" and r[2] == "-":
            line, src = int(r[0]), r[1].strip()
            key = (cur, line)
            tot[key] = tot.get(key, 0) + int(r[7])
            samp[key] = samp.get(key, 0) + int(r[4])
            text[key] = src
        elif "CALL" in r[3] and int(r[7]) > 0 and (cur, line, r[2]) not in seen:
            seen.add((cur, line, r[2]))                       # inlined lines are listed once per file section
            calls.append((cur, line, int(r[7]), src[:80]))
    T, S = sum(tot.values()), sum(samp.values())
    print(f"warp instructions executed {T}   stall samples {S}")
    byfile = {}
    for (f, _), v in tot.items():
        byfile[f] = byfile.get(f, 0) + v
    for f, v in sorted(byfile.items(), key=lambda kv: -kv[1]):
        sf = sum(s for (ff, _), s in samp.items() if ff == f)
        print(f"  {f:28s} {v:12d} {100 * v / T:5.1f}% instr  {100 * sf / S:5.1f}% samples")
    print("calls executed (warp level):")
    for c in calls:
        print("  ", c)
    print(f"top {top} lines:")
    for key, v in sorted(tot.items(), key=lambda kv: -kv[1])[:top]:
        print(f"  {key[0]}:{key[1]:<4d} {100 * v / T:5.1f}% instr {100 * samp[key] / S:5.1f}% samples | {text[key][:90]}")


if __name__ == "__main__":
    main()
