#!/bin/bash
# round 2, run 4 (2 GPUs): cfg2 at N=2 (exchange breakdown + cross-check), BASELINE config 3 (256 x 1024 over 2 learners),
# and the trainer pipeline (cfg5 shape, reduced) with 2 learners on 2 GPUs
mkdir -p gpurun_out
L=gpurun_out/r2_run04.log
: > $L
nvidia-smi topo -m >> $L 2>&1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
echo "== cfg2 N=2" >> $L
timeout 600 $TR bench.py --gpus 2 --steps 4 --warmup 3 > gpurun_out/r2_run04_bench_cfg2_n2.json 2>> $L
echo "exit $?" >> $L
echo "== cfg3 N=2" >> $L
timeout 900 $TR bench.py --gpus 2 --config cfg3 --steps 2 --warmup 3 --lean > gpurun_out/r2_run04_bench_cfg3_n2.json 2>> $L
echo "exit $?" >> $L
echo "== trainer pipeline, 2 stub generators + 2 learners (reduced cfg5: batch 32 x 8 candidates x 512 tokens)" >> $L
timeout 600 python -m distrl_llm_b200.train_distributed --model random:qwen2.5-7b --learner grpo --number_of_actors 2 --number_of_learners 2 \
  --batch_size 32 --learner_chunk_size 0 --num_candidates 8 --topk 8 --max_new_tokens 512 --max_lora_rank 16 --episodes 1 --eval_every 0 \
  --bench --max_steps 3 > gpurun_out/r2_run04_trainer_2learners.json 2>> $L
echo "exit $?" >> $L
for f in cfg2_n2 cfg3_n2; do
python - $f >> $L 2>&1 <<'PY'
import json, sys
try:
    d = json.loads(open(f"gpurun_out/r2_run04_bench_{sys.argv[1]}.json").read().strip().splitlines()[-1])
    print(sys.argv[1], "ms", round(d["ms_per_step"], 1), "tok/s", round(d["value"]), "e2e ms", round(d["e2e"]["ms_per_step"], 1), "frac_packed", d["step_roofline"]["frac_of_packed_roofline"])
    print(json.dumps(d["exchange"]))
except Exception as e:
    print(sys.argv[1], "no result", e)
PY
done
tail -c 1500 gpurun_out/r2_run04_trainer_2learners.json >> $L
grep -v "^$" $L | tail -30
