run() { tag=$1; shift; env "$@" python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/s_$tag.log 2>&1; echo "$tag: $(tail -c 2500 gpurun_out/s_$tag.log | grep -o '"decoder_advance": [0-9.]*') $(tail -c 2500 gpurun_out/s_$tag.log | grep -o '"expand": [0-9.]*') $(tail -c 2500 gpurun_out/s_$tag.log | grep -o '"in_shared_memory": [0-9.]*')"; tail -1 gpurun_out/s_$tag.log | cut -c1-200 | grep -i error; }
run c4096 B2K_DEC_RS_CAPS=4096,4096,4096
run c3072 B2K_DEC_RS_CAPS=3072,3072,2048
run c2048 B2K_DEC_RS_CAPS=2048,2048,2048
true
