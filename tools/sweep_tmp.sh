run() { tag=$1; shift; env "$@" python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/s_$tag.log 2>&1; echo "$tag: $(tail -c 2500 gpurun_out/s_$tag.log | grep -o '"stage_ms": {[^}]*}') $(tail -c 2500 gpurun_out/s_$tag.log | grep -o '"eps_replay": [0-9.]*') $(tail -c 2500 gpurun_out/s_$tag.log | grep -o '"in_shared_memory": [0-9.]*')"; tail -1 gpurun_out/s_$tag.log | cut -c1-200 | grep -i error; }
run t512big B2K_DEC_THREADS=512 B2K_DEC_RS_CAPS=4096,4096,4096
run t512mid B2K_DEC_THREADS=512 B2K_DEC_RS_CAPS=2048,2048,2048
run t256big B2K_DEC_THREADS=256 B2K_DEC_RS_CAPS=2048,2048,2048
true
