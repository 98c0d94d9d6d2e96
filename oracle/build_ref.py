"""Recipe for oracle/_ref: a git-ignored copy of the reference files that hold 100 % of the path's arithmetic
(stylegan2/models.py, stylegan2/modules.py, stylegan2/utils.py, clip/model.py; gpt2/{model,sample,config,utils}.py for
the img2txt workload), so that ``bench.py --impl reference``
can time the reference's OWN modules on the GPU box (which has no /root/reference).  Run in the build container by
``__graft_entry__.build()``:

    python -m oracle.build_ref

Nothing under oracle/_ref is tracked by git (.gitignore) and nothing in it is read by the product path, the tests'
GPU parity checks or smoke(): it is the CPU baseline arm only.  The two ``__init__.py`` files written here are empty
package markers (the reference's stylegan2/__init__.py imports its trainer, which is not on the path).
"""
import os
import shutil
import sys

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")
FILES = ["stylegan2/models.py", "stylegan2/modules.py", "stylegan2/utils.py", "clip/model.py",
         "gpt2/model.py", "gpt2/sample.py", "gpt2/config.py", "gpt2/utils.py"]       # img2txt arm (config 5)


def main() -> int:
    if not os.path.isdir(REF):
        print(f"build_ref: {REF} is absent (GPU box): keeping whatever oracle/_ref holds")
        return 0
    for rel in FILES:
        dst = os.path.join(OUT, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(os.path.join(REF, rel), dst)
    for pkg in ("stylegan2", "clip", "gpt2"):
        open(os.path.join(OUT, pkg, "__init__.py"), "w").close()
    print(f"build_ref: copied {len(FILES)} reference files into {OUT}")
    return 0


if __name__ == "__main__":
    sys.exit(main())
