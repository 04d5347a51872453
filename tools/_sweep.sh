for s in 4 8 2; do echo -n "side=$s "; PLK_MSM_SIDE_STREAMS=$s python tools/prover_mix.py --log-n 16 --reps 10 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print(d['ms_per_proof_mix'])"; done
for s in 4 8; do echo -n "lg18 side=$s "; PLK_MSM_SIDE_STREAMS=$s python tools/prover_mix.py --log-n 18 --reps 5 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print(d['ms_per_proof_mix'])"; done
