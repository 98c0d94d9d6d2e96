# A/B of two builds inside one run: this tree vs a variant library (CLIPGLASS_LIB), per-layer breakdown of selected layers
cd $GRAFT_REPO_ROOT
TAG=${1:-r02ab2}; VAR=${2:-clip_glass_b200/libclipglass_b200_regpf.so}; PAT=${3:-^G16}
for round in 1 2 3; do
  timeout 300 python tests/profile_step.py --pop 64 --evals 4 --timing 2>&1 | grep -E "$PAT|total conv" | sed "s/^/A (this tree) /"
  CLIPGLASS_LIB=$VAR timeout 300 python tests/profile_step.py --pop 64 --evals 4 --timing 2>&1 | grep -E "$PAT|total conv" | sed "s/^/B (variant)   /"
done | tee gpurun_out/ab2_$TAG.log
