timeout 600 python -m pytest tests/test_gpu_batch.py -m gpu -x -q 2>&1 | tail -2
for lib in "" u4m6 u4m5 u4m4 u3m5; do
    L=""; if [ -n "$lib" ]; then L=$PWD/multicharge_b200/libmchrg_b200_$lib.so; fi
    MCHRG_B200_LIB=$L timeout 300 python bench.py --steps 3 --warmup 3 --no-extra --cpu-sample 2000 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('lib=$lib', round(d['value']), round(d['e2e']['value']), 'factor ms', round(1e3*d['roofline']['kernel_seconds_per_step'],2), 'frac', round(d['roofline']['frac'],3), 'failed', d['failed_molecules'])"
done
