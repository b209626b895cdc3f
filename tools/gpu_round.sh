#!/bin/bash
# One gpurun call for a whole validation pass (box acquisition + snapshot push cost 1-3 GPU-minutes per call, so batch):
#   /usr/local/graft/bin/gpurun --timeout 900 -- 'bash tools/gpu_round.sh [tests] [bench] [ops] [launches] [full]'
# Each stage has its own timeout so a hung kernel cannot hold the box; outputs land in gpurun_out/.
mkdir -p gpurun_out
stages="${*:-tests bench}"
for s in $stages; do
  case "$s" in
    tests)    timeout 900 python -m pytest tests -q -m gpu --timeout 200 --tb=short -x --maxfail=25 > gpurun_out/pytest_gpu.log 2>&1; grep -v "Warning" gpurun_out/pytest_gpu.log | tail -60 ;;
    testsall) timeout 1200 python -m pytest tests -q -m gpu --timeout 200 --tb=short --maxfail=40 > gpurun_out/pytest_gpu.log 2>&1; grep -v "Warning" gpurun_out/pytest_gpu.log | tail -150
              if ! tail -1 gpurun_out/pytest_gpu.log | grep -q " passed" || tail -1 gpurun_out/pytest_gpu.log | grep -q failed; then
                echo "== bisect: scorer tests with the regular last-chunk mapping (PTRB200_NO_PARTIAL=1)"
                PTRB200_NO_PARTIAL=1 PTRB200_NO_RUNS=1 timeout 600 python -m pytest tests/test_gpu_scorer.py tests/test_gpu_tc_gemm.py tests/test_gpu_r2_parity.py tests/test_gpu_losses.py -q -m gpu --timeout 200 --tb=line 2>&1 | grep -v Warning | tail -15
              fi ;;
    smoke)    timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 ;;
    bench)    timeout 200 python bench.py 2>gpurun_out/bench_n1.err | tail -1 > gpurun_out/bench_n1.json
              python -c "import json; d=json.loads(open('gpurun_out/bench_n1.json').read()); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['gpu_launches']); print(d['roofline']['kernels_ms_per_step'])" ;;
    ops)      timeout 300 python tools/op_bench.py > gpurun_out/op_bench.log 2>&1; cp profiles/r02_op_table.md gpurun_out/; tail -60 gpurun_out/r02_op_table.md ;;
    launches) timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv python tools/profile_step.py 3 > gpurun_out/ncu_launch.log 2>&1; tail -1 gpurun_out/ncu_launch.log ;;
    quick)    # headline + list-scorer step, kernel shares only
              for c in b c; do
                timeout 300 python bench.py --config $c --steps 50 --no-cpu-baseline 2>gpurun_out/bench_q_$c.err | tail -1 > gpurun_out/bench_q_$c.json
                python -c "import json; d=json.loads(open('gpurun_out/bench_q_$c.json').read()); print('$c', round(d['value'],1), 'q/s', round(d['ms_per_step'],4), 'ms', {k:v for k,v in list(d['roofline']['kernels_ms_per_step'].items())[:8]})" || tail -3 gpurun_out/bench_q_$c.err
              done ;;
    abwg)     # A/B of the deeper TMA ring (shorter row tiles) in the weight-gradient kernel
              for v in 0 1; do
                if [ $v = 1 ]; then export PTRB200_WG_DEEP=1; else unset PTRB200_WG_DEEP; fi
                for c in b c; do
                  timeout 300 python bench.py --config $c --steps 50 --no-cpu-baseline 2>/dev/null | tail -1 > gpurun_out/bench_abwg_${c}_$v.json
                  python -c "import json; d=json.loads(open('gpurun_out/bench_abwg_${c}_$v.json').read()); print('DEEP=$v', '$c', round(d['ms_per_step'],4), 'wgrad', d['roofline']['kernels_ms_per_step'].get('wgrad_tc'))"
                done
              done; unset PTRB200_WG_DEEP ;;
    multi8)   # 8 ranks: headline bench with the peer-memory exchange and with NCCL, config d (LambdaLoss n=1024), the reference arm
              for v in peer nccl; do
                if [ $v = nccl ]; then export PTRANKING_B200_PEER=0; else unset PTRANKING_B200_PEER; fi
                timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 8 --steps 100 --warmup 3 2>gpurun_out/bench_n8_$v.err | tail -1 > gpurun_out/bench_n8_$v.json
                python -c "import json; d=json.loads(open('gpurun_out/bench_n8_$v.json').read()); print('N=8 $v', d['config'].get('gradient_exchange')[:40], round(d['value'],1), round(d['ms_per_step'],4), 'strong', round(d['strong_scaling']['ms_per_step'],4), 'e2e', round(d['e2e']['value'],1))" || tail -5 gpurun_out/bench_n8_$v.err
              done; unset PTRANKING_B200_PEER
              timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29532 bench.py --config d --gpus 8 --steps 50 --warmup 3 2>gpurun_out/bench_d_n8.err | tail -1 > gpurun_out/bench_d_n8.json
              python -c "import json; d=json.loads(open('gpurun_out/bench_d_n8.json').read()); print('N=8 config d', round(d['value'],1), round(d['ms_per_step'],4))" || tail -5 gpurun_out/bench_d_n8.err
              timeout 300 python bench.py --steps 100 --no-cpu-baseline 2>/dev/null | tail -1 > gpurun_out/bench_n1_same_box.json
              python -c "import json; d=json.loads(open('gpurun_out/bench_n1_same_box.json').read()); print('N=1 same box', round(d['value'],1), round(d['ms_per_step'],4))" ;;
    abcs)     # A/B of the next-row prefetch in the dY / statistics sweep
              for v in 0 1; do
                if [ $v = 1 ]; then export PTRB200_NO_CS_PREFETCH=1; else unset PTRB200_NO_CS_PREFETCH; fi
                timeout 300 python bench.py --steps 50 --no-cpu-baseline 2>/dev/null | tail -1 > gpurun_out/bench_abcs_$v.json
                python -c "import json; d=json.loads(open('gpurun_out/bench_abcs_$v.json').read()); print('NO_CS_PREFETCH=$v', round(d['ms_per_step'],4), 'colstat_dy', d['roofline']['kernels_ms_per_step'].get('colstat_dy'))"
              done; unset PTRB200_NO_CS_PREFETCH ;;
    abc)      # config c (list scorer, L=6): aligned bgemm kernel and fused Q|K|V projection, each switched off in turn
              for v in base general_bgemm separate_qkv; do
                unset PTRB200_BGEMM_GENERAL PTRANKING_B200_FUSED_QKV
                [ $v = general_bgemm ] && export PTRB200_BGEMM_GENERAL=1
                [ $v = separate_qkv ] && export PTRANKING_B200_FUSED_QKV=0
                timeout 300 python bench.py --config c --no-cpu-baseline 2>gpurun_out/bench_c_$v.err | tail -1 > gpurun_out/bench_c_$v.json
                python -c "import json; d=json.loads(open('gpurun_out/bench_c_$v.json').read()); print('$v', round(d['value'],1), 'q/s', round(d['ms_per_step'],4), 'ms', {k:v for k,v in list(d['roofline']['kernels_ms_per_step'].items())[:12]})" || tail -3 gpurun_out/bench_c_$v.err
              done; unset PTRB200_BGEMM_GENERAL PTRANKING_B200_FUSED_QKV ;;
    launchesc) # per-launch list of one list-scorer step (config c, L=3)
              ENC_LAYERS=3 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/launches_c.csv python tools/profile_step.py 2 0 c > gpurun_out/ncu_launch_c.log 2>&1; tail -1 gpurun_out/ncu_launch_c.log
              python tools/launch_table.py gpurun_out/launches_c.csv > gpurun_out/launch_table_c.txt; head -1 gpurun_out/launch_table_c.txt ;;
    full)     timeout 800 bash tools/profile_all.sh 2>&1 | tail -3 ;;
    ab)       # A/B of the short-last-chunk staging in the forward layer kernel
              for v in 0 1; do
                if [ $v = 1 ]; then export PTRB200_NO_PARTIAL=1; else unset PTRB200_NO_PARTIAL; fi
                timeout 200 python bench.py --steps 50 --no-cpu-baseline 2>/dev/null | tail -1 > gpurun_out/bench_ab_$v.json
                python -c "import json; d=json.loads(open('gpurun_out/bench_ab_$v.json').read()); print('NO_PARTIAL=$v', round(d['ms_per_step'],4), {k:v for k,v in list(d['roofline']['kernels_ms_per_step'].items())[:6]})"
              done; unset PTRB200_NO_PARTIAL ;;
    configs)  for c in a c d e; do
                timeout 400 python bench.py --config $c 2>gpurun_out/bench_$c.err | tail -1 > gpurun_out/bench_$c.json
                python -c "import json; d=json.loads(open('gpurun_out/bench_$c.json').read()); print('$c', round(d['value'],1), 'q/s', round(d['ms_per_step'],4), 'ms; e2e', round(d['e2e']['value'],1), '; ref_cuda', d.get('reference_cuda'), '; cpu', d.get('cpu_baseline',{}).get('value'), d['roofline'].get('frac'), d['roofline'].get('step_frac'))" || tail -3 gpurun_out/bench_$c.err
              done
              timeout 400 python bench.py --config c --enc-layers 3 2>/dev/null | tail -1 > gpurun_out/bench_c_L3.json
              for n in 32 1024; do timeout 300 python bench.py --config e --docs $n --no-cpu-baseline 2>/dev/null | tail -1 > gpurun_out/bench_e_n$n.json; done
              python -c "import json; [print(f, round(json.loads(open('gpurun_out/'+f).read())['value'],1)) for f in ('bench_c_L3.json','bench_e_n32.json','bench_e_n1024.json')]" ;;
    parity)   timeout 400 python tools/parity_table.py r02 > gpurun_out/parity_table.log 2>&1; tail -5 gpurun_out/parity_table.log; cp profiles/r02_parity_table.md gpurun_out/ 2>/dev/null ;;
    dropin)   timeout 600 python tools/dropin_run.py --impl b200 --model LambdaRank > gpurun_out/dropin_run.log 2>&1; grep -v Warning gpurun_out/dropin_run.log | tail -14
              timeout 600 python tools/dropin_run.py --impl b200 --model ApproxNDCG --sf listsf --queries 120 > gpurun_out/dropin_run_listsf.log 2>&1; grep -v Warning gpurun_out/dropin_run_listsf.log | tail -8 ;;
    benchA)   timeout 400 python bench.py --config a 2>gpurun_out/bench_a.err | tail -1 > gpurun_out/bench_a.json
              python -c "import json; d=json.loads(open('gpurun_out/bench_a.json').read()); print('a', round(d['value'],1), round(d['ms_per_step'],4), d.get('reference_cuda'), list(d['roofline']['kernels_ms_per_step'].items())[:6])" ;;
    benchDE)  for c in d e; do timeout 400 python bench.py --config $c --no-cpu-baseline 2>/dev/null | tail -1 > gpurun_out/bench_${c}_nocpu.json; done ;;
    profile)  timeout 1500 bash tools/profile_r02.sh launches b c d e 2>&1 | tail -20 ;;
    abkt)     # A/B of the width-specialised layer kernels
              for v in 0 1; do
                if [ $v = 1 ]; then export PTRB200_NO_KT=1; else unset PTRB200_NO_KT; fi
                timeout 200 python bench.py --steps 50 --no-cpu-baseline 2>/dev/null | tail -1 > gpurun_out/bench_abkt_$v.json
                python -c "import json; d=json.loads(open('gpurun_out/bench_abkt_$v.json').read()); print('NO_KT=$v', round(d['ms_per_step'],4), {k:v for k,v in list(d['roofline']['kernels_ms_per_step'].items())[:6]})"
              done; unset PTRB200_NO_KT ;;
    multi)    # two ranks: device-side data-parallel test + the N=2 bench line (weak + strong scaling)
              timeout 900 python -m pytest tests/test_gpu_multirank.py -m gpu -q -x --timeout 400 --tb=short > gpurun_out/multirank_test.log 2>&1; grep -v Warning gpurun_out/multirank_test.log | tail -25
              timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 100 --warmup 3 2>gpurun_out/bench_n2.err | tail -1 > gpurun_out/bench_n2.json
              python -c "import json; d=json.loads(open('gpurun_out/bench_n2.json').read()); print('N=2', d['config'].get('gradient_exchange'), round(d['value'],1), round(d['ms_per_step'],4), 'strong', round(d['strong_scaling']['ms_per_step'],4))" || tail -5 gpurun_out/bench_n2.err
              PTRANKING_B200_PEER=0 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29520 bench.py --gpus 2 --steps 100 --warmup 3 2>/dev/null | tail -1 > gpurun_out/bench_n2_nccl.json
              python -c "import json; d=json.loads(open('gpurun_out/bench_n2_nccl.json').read()); print('N=2 NCCL overlapped', round(d['value'],1), round(d['ms_per_step'],4), 'strong', round(d['strong_scaling']['ms_per_step'],4))"
              PTRANKING_B200_PEER=0 PTRANKING_B200_OVERLAP=0 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 2 --steps 100 --warmup 3 2>/dev/null | tail -1 > gpurun_out/bench_n2_nooverlap.json
              python -c "import json; d=json.loads(open('gpurun_out/bench_n2_nooverlap.json').read()); print('N=2 NCCL no overlap', round(d['value'],1), round(d['ms_per_step'],4), 'strong', round(d['strong_scaling']['ms_per_step'],4))"
              timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29519 bench.py --impl reference --gpus 2 --steps 5 --warmup 3 2>/dev/null | tail -1 | cut -c1-300 ;;
    *)        echo "unknown stage $s" ;;
  esac
done
