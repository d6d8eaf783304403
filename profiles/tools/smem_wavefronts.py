"""Shared-memory wavefronts (the LSU data-pipe unit, 128 B) per source line of one kernel, from an ncu report captured
with --set full --import-source on.   python profiles/tools/smem_wavefronts.py report.ncu-rep <kernel substring> [top_n]
Also prints SHFL instructions (they use the same pipe) and the totals per frame when the kernel calls warp_frame_mfcc."""
import csv
import subprocess
import sys


def main():
    rep, kern = sys.argv[1], sys.argv[2]
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, inside, taken = None, False, False
    cur = line = src = None
    wf, ex, ins, shfl, text = {}, {}, {}, {}, {}
    for r in rows:
        if len(r) >= 2 and r[0] == "Function Name":
            if taken and inside:
                break                                   # first matching launch only
            inside = kern in r[1]
            taken = taken or inside
            continue
        if not inside:
            continue
        if len(r) >= 2 and r[0] == "File Path":
            cur = r[1].split("/")[-1]
            continue
        if r and r[0] == "Line No":
            hdr = {n: i for i, n in enumerate(r)}
            continue
        if hdr is None or len(r) < len(hdr) - 5:
            continue
        if r[0] != "" and r[2] == "-":
            line, src = int(r[0]), r[1].strip()
            text[(cur, line)] = src
        elif r[0] == "" and line is not None:
            key = (cur, line)
            num = lambda c: int(r[hdr[c]]) if r[hdr[c]].isdigit() else 0
            w, e, n = num("L1 Wavefronts Shared"), num("L1 Wavefronts Shared Excessive"), num("Instructions Executed")
            wf[key] = wf.get(key, 0) + w
            ex[key] = ex.get(key, 0) + e
            ins[key] = ins.get(key, 0) + n
            if "SHFL" in r[3]:
                shfl[key] = shfl.get(key, 0) + n
    W, E, S, I = sum(wf.values()), sum(ex.values()), sum(shfl.values()), sum(ins.values())
    print(f"kernel {kern}: instructions {I}  shared wavefronts {W} (excessive {E})  SHFL {S}")
    byfile = {}
    for (f, _), v in wf.items():
        byfile[f] = byfile.get(f, 0) + v
    for f, v in sorted(byfile.items(), key=lambda kv: -kv[1]):
        if v:
            print(f"  {str(f):28s} {v:12d} {100 * v / max(1, W):5.1f}%")
    print(f"top {top} lines by shared wavefronts + SHFL:")
    keys = sorted(set(wf) | set(shfl), key=lambda k: -(wf.get(k, 0) + shfl.get(k, 0)))[:top]
    for k in keys:
        print(f"  {str(k[0])}:{k[1]:<4d} wavefronts {wf.get(k, 0):9d} (excess {ex.get(k, 0):8d})  shfl {shfl.get(k, 0):8d}  instr {ins.get(k, 0):9d} | {text.get(k, '')[:80]}")


if __name__ == "__main__":
    main()
