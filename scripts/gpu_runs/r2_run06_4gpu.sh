#!/bin/bash
# round 2, run 6 (4 GPUs): BASELINE config 4 (1024 x 2048 sampled, top-k 128 of 256 -> 512 sequences over 4 learners) and
# the trainer pipeline with 4 stub generators + 4 learners (config 5 shape; generators share the learners' GPUs)
mkdir -p gpurun_out
L=gpurun_out/r2_run06.log
: > $L
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29512"
echo "== cfg4 N=4" >> $L
timeout 1500 $TR bench.py --gpus 4 --config cfg4 --steps 1 --warmup 3 --lean > gpurun_out/r2_run06_bench_cfg4_n4.json 2>> $L
echo "exit $?" >> $L
python - >> $L 2>&1 <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/r2_run06_bench_cfg4_n4.json").read().strip().splitlines()[-1])
    print("cfg4 n4: ms", round(d["ms_per_step"], 1), "tok/s", round(d["value"]), "e2e ms", round(d["e2e"]["ms_per_step"], 1), "frac_packed", d["step_roofline"]["frac_of_packed_roofline"], d["config"]["passes"])
    print(json.dumps(d["exchange"]))
    print(d["profile_ms"])
except Exception as e:
    print("cfg4 no result", e)
PY
echo "== trainer pipeline, 4 stub generators + 4 learners: batch 128 x 16 candidates x 512 tokens (cfg5 shape, reduced 4x in batch and 2.3x in length)" >> $L
timeout 900 python -m distrl_llm_b200.train_distributed --model random:qwen2.5-7b --learner grpo --number_of_actors 4 --number_of_learners 4 \
  --batch_size 128 --learner_chunk_size 0 --num_candidates 16 --topk 16 --max_new_tokens 512 --max_lora_rank 16 --episodes 1 --eval_every 0 \
  --bench --max_steps 3 > gpurun_out/r2_run06_trainer_4learners.json 2>> $L
echo "exit $?" >> $L
tail -c 1500 gpurun_out/r2_run06_trainer_4learners.json >> $L
echo "== same with generation overlapped (N3)" >> $L
timeout 900 python -m distrl_llm_b200.train_distributed --model random:qwen2.5-7b --learner grpo --number_of_actors 4 --number_of_learners 4 \
  --batch_size 128 --learner_chunk_size 0 --num_candidates 16 --topk 16 --max_new_tokens 512 --max_lora_rank 16 --episodes 1 --eval_every 0 \
  --bench --max_steps 3 --overlap_generation > gpurun_out/r2_run06_trainer_4learners_overlap.json 2>> $L
echo "exit $?" >> $L
tail -c 1500 gpurun_out/r2_run06_trainer_4learners_overlap.json >> $L
grep -v "^$" $L | tail -25
