"""Dev tool: prove that a refactor left the DEFAULT kernels untouched at the machine-code level.

Compiles csrc/{logsig,logncde,linear_scan}.cu of a git revision and of the working tree for sm_100a and compares
the SASS instruction streams kernel by kernel (template arguments added since - the trailing `FUSED = false` of
the specialised Log-NCDE kernels - are normalised away; kernels that exist only in the working tree are listed).

    python scripts/sass_diff.py 4618d2f       (the sources measured on B200 in round 1)
"""
import os
import re
import subprocess
import sys
import tempfile

R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FILES = ("logsig", "logncde", "linear_scan")
NVCC = ["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17"]


def compile_tree(root, out):
    procs = []
    for f in FILES:
        procs.append(subprocess.Popen(NVCC + ["-I", f"{root}/include", "-I", f"{root}/log-neural-cdes_b200/csrc", "-c",
                                             f"{root}/log-neural-cdes_b200/csrc/{f}.cu", "-o", f"{out}/{f}.o"],
                                      stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL))
    assert all(p.wait() == 0 for p in procs)


def sass(obj):
    out = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
    funcs, cur = {}, None
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            funcs[cur] = []
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4}\*/\s+(.*?);", line)
        if m and cur:
            funcs[cur].append(re.sub(r"\s+", " ", m.group(1)))
    names = list(funcs)
    dem = subprocess.run(["c++filt"] + names, capture_output=True, text=True).stdout.splitlines()
    return {d.replace("lncde::fast::", "").replace("lncde::", ""): funcs[n] for n, d in zip(names, dem)}


def norm(name):
    m = re.match(r"void (logncde_fwd_fast_kernel)<(\d), (\w+), (\w+), (\w+)>", name)
    if m:
        return None if m.group(5) == "true" else f"void {m.group(1)}<{m.group(2)}, {m.group(3)}, {m.group(4)}>(FwdParams)"
    m = re.match(r"void (logncde_bwd_(?:fast|saved)_kernel)<(\d), (\w+), (\w+)(?:, (\w+))?>", name)
    if m:
        variant = m.group(4) == "true" or m.group(5) == "true"
        return None if variant else f"void {m.group(1)}<{m.group(2)}, {m.group(3)}>(BwdParams)"
    m = re.match(r"void (logncde_(?:fwd|bwd)_kernel)<(\d), (\w+)>\((\w+)\)", name)
    if m:  # generic kernels: trailing PF = false of the streamed-weights variant ("k3_stream")
        return None if m.group(3) == "true" else f"void {m.group(1)}<{m.group(2)}>({m.group(4)})"
    return name


def main(rev):
    with tempfile.TemporaryDirectory() as tmp:
        old_root = os.path.join(tmp, "old")
        os.makedirs(old_root)
        subprocess.run(["git", f"--work-tree={old_root}", "checkout", rev, "--", "log-neural-cdes_b200/csrc", "include"],
                       cwd=R, check=True)
        subprocess.run(["git", "reset", "-q"], cwd=R, check=True)
        os.makedirs(f"{tmp}/o"), os.makedirs(f"{tmp}/n")
        compile_tree(old_root, f"{tmp}/o")
        compile_tree(R, f"{tmp}/n")
        bad = 0
        for f in FILES:
            old = sass(f"{tmp}/o/{f}.o")
            new = {norm(k): v for k, v in sass(f"{tmp}/n/{f}.o").items() if norm(k)}
            same = [k for k in old if old[k] == new.get(k)]
            diff = [k for k in old if k in new and old[k] != new[k]]
            gone = [k for k in old if k not in new]
            print(f"{f}.cu: {len(same)} kernels identical, {len(diff)} different, {len(gone)} removed, "
                  f"{len([k for k in new if k not in old])} new")
            for k in diff + gone:
                print("   ", k)
            bad += len(diff) + len(gone)
        return bad


if __name__ == "__main__":
    sys.exit(1 if main(sys.argv[1] if len(sys.argv) > 1 else "4618d2f") else 0)
