O=gpurun_out/r2q; mkdir -p $O; cd /root/repo
B2K_DEC_PROF=1 timeout 300 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --parity-utts 0 > $O/bench_prof.json 2> $O/bench_prof.err
B2K_DEC_PREFETCH=1 B2K_DEC_PARWALK=0 timeout 300 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --parity-utts 1 > $O/bench_pf.json 2> $O/bench_pf.err
B2K_DEC_PARWALK=0 timeout 300 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --parity-utts 0 > $O/bench_nopar.json 2> $O/bench_nopar.err
for f in prof pf nopar; do python -c "
import json,sys
d=json.loads(open('$O/bench_$f.json').read().strip().splitlines()[-1]); print('$f', round(d['value']), d['stage_ms'], d.get('parity_checked')); print(d.get('decoder_phase_share'), d.get('eps_replay_per_frame'), d.get('replay_routes'))"; tail -n 3 $O/bench_$f.err; done
