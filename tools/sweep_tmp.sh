run() { python bench.py --steps 5 --warmup 3 --no-e2e 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); k = d['kernels_ms_per_step']
        print('$1', d['ms_per_step'], {n: v for n, v in k.items() if 'mu_t' in n or 'advance_uv' in n}, d['checksum']['u_2'])
"; }

run default
for v in s4 s5b3 s6b3 cw16s4; do WRFGPU_LIB=$PWD/wrf_b200/variants/libwrfgpu_$v.so run $v; done
