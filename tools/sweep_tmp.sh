run() { tag=$1; shift; env "$@" python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/s_$tag.log 2>&1; echo "$tag: $(tail -c 2500 gpurun_out/s_$tag.log | grep -o '"stage_ms": {[^}]*}') $(tail -c 2500 gpurun_out/s_$tag.log | grep -o '"eps_replay": [0-9.]*')"; }
run base A=1
run prefetch B2K_DEC_TUNE=1
run fin512 B2K_FIN_THREADS=512
run fin1024 B2K_FIN_THREADS=1024
