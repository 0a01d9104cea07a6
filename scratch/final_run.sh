#!/bin/bash
# final validation of a round: GPU tests, reference arm, bench, ncu launch list, ncu full capture of the dominant kernel
tag=$1
python -m pytest tests -m gpu -q > gpurun_out/${tag}_pytest.log 2>&1; tail -2 gpurun_out/${tag}_pytest.log
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/${tag}_bench_ref.json 2> gpurun_out/${tag}_bench_ref.err
python bench.py > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; tail -c 300 gpurun_out/${tag}_bench.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/${tag}_ncu_l.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_lp_grad_mom --launch-skip 4 -c 1 -o gpurun_out/${tag}_lpgrad -f python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/${tag}_ncu_f.log 2>&1
python -c "import __graft_entry__ as g; g.smoke()"
