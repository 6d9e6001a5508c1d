export B200RL_PROFILE_ONE_STEP=1
timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct --clock-control none --profile-from-start off -k regex:"gemm_pair_kernel" -s 108 -c 10 --csv --log-file gpurun_out/r2_run27_gemm_traffic_1pass.csv python bench.py --steps 1 --warmup 1 --no_cpu_baseline > /dev/null 2>&1
echo "exit $?"
python - <<'PY'
import csv,io
txt=open('gpurun_out/r2_run27_gemm_traffic_1pass.csv').read()
start=txt.index('"ID"')
by={}
for x in csv.DictReader(io.StringIO(txt[start:])):
    by.setdefault(int(x['ID']),{})[x['Metric Name']]=float(x['Metric Value'].replace(',',''))
for i in sorted(by):
    m=by[i]; print(i, 'rd %.3f GB wr %.3f GB hit %.1f%% %.3f ms'%(m['dram__bytes_read.sum']/1e9, m['dram__bytes_write.sum']/1e9, m['lts__t_sector_hit_rate.pct'], m['gpu__time_duration.sum']/1e6))
PY
