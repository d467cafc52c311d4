# usage: tools/ncu_front.sh <tag> [frames]   -> gpurun_out/<tag>.ncu-rep (one front-kernel launch, --set full, source import)
tag=$1; frames=${2:-9}
timeout 600 ncu --set full --import-source on --clock-control none -k regex:front -c 1 -f -o gpurun_out/$tag python bench.py --frames $frames --steps 1 --warmup 1 --no-e2e --no-cpu --no-bands --no-extras > gpurun_out/$tag.log 2>&1
tail -2 gpurun_out/$tag.log
