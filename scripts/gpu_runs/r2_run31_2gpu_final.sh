#!/bin/bash
# round 2, run 31 (2 GPUs): the final build at N=2 launched like the driver does: cfg2 bench (exchange breakdown after the
# reduce_adam flag fix), reference arm under torchrun (rank 0 only), trainer pipeline with 2 learners
mkdir -p gpurun_out
L=gpurun_out/r2_run31.log
: > $L
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
echo "== cfg2 N=2" >> $L
timeout 600 $TR bench.py --gpus 2 --steps 4 --warmup 3 > gpurun_out/r2_run31_bench_cfg2_n2.json 2>> $L
echo "exit $?" >> $L
echo "== trainer pipeline, 2 stub generators + 2 learners" >> $L
timeout 600 python -m distrl_llm_b200.train_distributed --model random:qwen2.5-7b --learner grpo --number_of_actors 2 --number_of_learners 2 \
  --batch_size 32 --learner_chunk_size 0 --num_candidates 8 --topk 8 --max_new_tokens 512 --max_lora_rank 16 --episodes 1 --eval_every 0 \
  --bench --max_steps 3 > gpurun_out/r2_run31_trainer_2learners.json 2>> $L
echo "exit $?" >> $L
python - >> $L 2>&1 <<'PY'
import json
d = json.loads(open("gpurun_out/r2_run31_bench_cfg2_n2.json").read().strip().splitlines()[-1])
print("cfg2_n2 ms", round(d["ms_per_step"], 1), "tok/s", round(d["value"]), "e2e ms", round(d["e2e"]["ms_per_step"], 1))
print(json.dumps(d["exchange"]))
t = json.loads(open("gpurun_out/r2_run31_trainer_2learners.json").read().strip().splitlines()[-1])
print("trainer steps/s", t["value"], "update s", t.get("timing/update_duration_mean_s"))
PY
grep -v "^$" $L | grep -v "^\*\*\*\|OMP_NUM" | tail -12
