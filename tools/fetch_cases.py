"""Copies the reference's INPUT DATA (case configs, STL geometry, golden logs/CSVs — no source code) from
/root/reference into baseline/_ref/ so that it travels to the GPU box with the working tree (git-ignored)."""
import os, shutil, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = "/root/reference"
def main():
    if not os.path.isdir(SRC):
        return False
    dst = os.path.join(ROOT, "baseline", "_ref")
    os.makedirs(os.path.join(dst, "CASES"), exist_ok=True)
    for case in sorted(os.listdir(os.path.join(SRC, "CASES"))):
        s = os.path.join(SRC, "CASES", case)
        if not os.path.isdir(s):
            continue
        d = os.path.join(dst, "CASES", case)
        os.makedirs(d, exist_ok=True)
        for f in os.listdir(s):
            if f.endswith((".yaml", ".stl")):
                shutil.copy2(os.path.join(s, f), os.path.join(d, f))
        if os.path.isdir(os.path.join(s, "RESULTS")):
            shutil.copytree(os.path.join(s, "RESULTS"), os.path.join(d, "RESULTS"), dirs_exist_ok=True)
    for f in os.listdir(SRC):
        if f.startswith("RESULTS_") and f.endswith(".txt"):
            shutil.copy2(os.path.join(SRC, f), os.path.join(dst, f))
    return True
if __name__ == "__main__":
    print("copied" if main() else "no /root/reference here")
