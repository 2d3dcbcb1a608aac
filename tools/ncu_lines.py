"""Per-source-line instruction and stall-sample totals of one kernel from an ncu report (no GPU needed):

    python tools/ncu_lines.py gpurun_out/prof.ncu-rep bucket_join cloudbrush_b200/libcloudbrush_gpu.so [top]

`ncu --page source --csv` lists SASS instructions with their counters; `nvdisasm -g` gives the source
line of every SASS offset of the same kernel (the .so must be the build that was profiled, -lineinfo)."""
import csv
import os
import re
import subprocess
import sys
import tempfile


def sass_lines(so, kernel_pat):
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(so)], cwd=tmp, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    for f in sorted(os.listdir(tmp)):
        if not f.endswith(".cubin"):
            continue
        txt = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, f)], capture_output=True, text=True).stdout
        cur, line, inl, res = None, None, None, {}
        for l in txt.split("\n"):
            m = re.match(r"\s*\.text\.(\S+):", l)
            if m:
                cur = m.group(1)
                res.setdefault(cur, [])
                continue
            m = re.search(r'//## File "([^"]+)", line (\d+)(?: inlined at "([^"]+)", line (\d+))?', l)
            if m:
                line = (os.path.basename(m.group(1)), int(m.group(2)))
                inl = (os.path.basename(m.group(3)), int(m.group(4))) if m.group(3) else None
                continue
            m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
            if m and cur:
                res[cur].append((int(m.group(1), 16), line, inl, m.group(2)))
        for name, ins in res.items():
            if re.search(kernel_pat, name) and ins:
                yield name, ins


def main():
    rep, pat, so = sys.argv[1], sys.argv[2], sys.argv[3]
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 60
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + pat],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(out.split("\n")))
    kernels, cur = [], None
    for r in rows:
        if r and r[0] == "Kernel Name":
            cur = {"name": r[1], "hdr": None, "ins": []}
            kernels.append(cur)
        elif r and r[0] == "Address" and cur is not None:
            cur["hdr"] = r
        elif cur is not None and cur["hdr"] and len(r) > 8:
            cur["ins"].append(dict(zip(cur["hdr"], r)))
    k = kernels[0]
    base = int(k["ins"][0]["Address"], 16)
    cands = list(sass_lines(so, pat))
    name, ins = [c for c in cands if len(c[1]) == len(k["ins"])][0] if any(len(c[1]) == len(k["ins"]) for c in cands) else cands[0]
    off2line = {o: (ln, inl) for o, ln, inl, _ in ins}
    agg, tot_i, tot_s = {}, 0, 0
    for d in k["ins"]:
        off = int(d["Address"], 16) - base
        ln, inl = off2line.get(off, (None, None))
        key = inl if (inl and "stage_" in inl[0]) else ln  # attribute inlined helpers to the call site in the stage file
        ie = int(d["Instructions Executed"] or 0)
        sm = int(d["# Samples"] or 0)
        te = int(d["Thread Instructions Executed"] or 0)
        a = agg.setdefault(key, [0, 0, 0])
        a[0] += ie
        a[1] += sm
        a[2] += te
        tot_i += ie
        tot_s += sm
    print(f"{k['name']}: {tot_i} warp instructions, {tot_s} samples, {len(k['ins'])} SASS instructions")
    src = {}
    for key, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
        text = ""
        if key:
            path = os.path.join(os.path.dirname(os.path.abspath(so)), "csrc", key[0])
            if path not in src and os.path.exists(path):
                src[path] = open(path).read().split("\n")
            if path in src and key[1] - 1 < len(src[path]):
                text = src[path][key[1] - 1].strip()[:100]
        print(f"{str(key):34s} instr {v[0] / tot_i * 100:5.1f}%  samples {v[1] / max(tot_s, 1) * 100:5.1f}%  thr/warp {v[2] / max(v[0], 1):4.1f}  {text}")


if __name__ == "__main__":
    main()
