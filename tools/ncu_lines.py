"""Attribute ncu per-SASS-instruction counters to CUDA source lines (dev tool).

usage: ncu_lines.py <report.ncu-rep> <lib.so> <kernel-name-substring> [top]
Needs ncu, cuobjdump and nvdisasm on PATH; the library must be built with -lineinfo.
"""
import csv
import glob
import io
import os
import re
import subprocess
import sys
import tempfile


def sass_lines(lib, kernel_sub):
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=tmp, capture_output=True)
    out = {}
    for cubin in glob.glob(os.path.join(tmp, "*.cubin")):
        text = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout
        cur_fn, cur_line = None, None
        for ln in text.splitlines():
            m = re.match(r"\s*\.section\s+\.text\.(\S+?),", ln)
            if m:
                cur_fn = m.group(1)
                continue
            m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
            if m:
                cur_line = (os.path.basename(m.group(1)), int(m.group(2)))
                continue
            m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
            if m and cur_fn:
                out.setdefault(cur_fn, {})[int(m.group(1), 16)] = (cur_line, m.group(2))
    return out


def sections(rep):
    text = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(text)))
    out, cur = [], None
    for r in rows:
        if r and r[0] == "Kernel Name":
            cur = {"name": r[1], "hdr": None, "rows": []}
            out.append(cur)
        elif r and r[0] == "Address" and cur is not None:
            cur["hdr"] = r
        elif cur is not None and cur["hdr"] and len(r) >= len(cur["hdr"]) - 2:
            cur["rows"].append(dict(zip(cur["hdr"], r)))
    return out


def match_fn(fn_maps, sec):
    """Pick the disassembled function whose first instructions equal the profiled ones."""
    base = int(sec["rows"][0]["Address"], 16)
    want = [(int(r["Address"], 16) - base, r["Source"].split()[0] if r["Source"].split() else "") for r in sec["rows"][:40]]
    best, best_score = None, -1
    for name, m in fn_maps.items():
        score = sum(1 for off, op in want if off in m and m[off][1].split() and
                    (m[off][1].split()[0] == op or (m[off][1].split()[0].startswith("@") and len(m[off][1].split()) > 1 and m[off][1].split()[1] == op)
                     or (op.startswith("@"))))
        if score > best_score and len(m) == len(sec["rows"]):
            best, best_score = name, score
    if best is None:
        for name, m in fn_maps.items():
            if len(m) == len(sec["rows"]):
                best = name
    return best


def main():
    rep, lib = sys.argv[1:3]
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
    fn_maps = sass_lines(lib, "")
    src_cache = {}
    for sec in sections(rep):
        if not sec["rows"]:
            continue
        key = match_fn(fn_maps, sec)
        base = int(sec["rows"][0]["Address"], 16)
        per_line, ti, ts = {}, 0, 0
        for d in sec["rows"]:
            off = int(d["Address"], 16) - base
            inst = int(d.get("Instructions Executed") or 0)
            samp = int(d.get("# Samples") or 0)
            line = fn_maps.get(key, {}).get(off, (None, ""))[0] if key else None
            e = per_line.setdefault(line, [0, 0])
            e[0] += inst
            e[1] += samp
            ti += inst
            ts += samp
        print(f"\n=== {sec['name'][:90]}\n    total warp-instructions {ti}  stall samples {ts}")
        for line, (inst, samp) in sorted(per_line.items(), key=lambda kv: -kv[1][0])[:top]:
            txt = ""
            if line and line[0]:
                path = glob.glob(os.path.join(os.path.dirname(os.path.abspath(lib)), line[0]))
                if path:
                    src_cache.setdefault(path[0], open(path[0]).read().splitlines())
                    if line[1] - 1 < len(src_cache[path[0]]):
                        txt = src_cache[path[0]][line[1] - 1].strip()
            print(f"{inst / max(ti, 1) * 100:5.1f}% inst {samp / max(ts, 1) * 100:5.1f}% stall  {str(line):28s} {txt[:95]}")


if __name__ == "__main__":
    main()
