timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fr.py tests/test_gpu_nsfr.py tests/test_gpu_scale.py tests/test_gmres.py tests/test_gpu_viscous.py -m gpu -x -q 2>&1 | tail -3
timeout 200 python tools/time_frjac.py 2>&1 | tail -1 | python -c "import sys,json; j=json.loads(sys.stdin.read()); print({k:v for k,v in j['ms'].items() if 'jac' in k or 'lu' in k or k=='total'})"
timeout 200 python tools/time_pgjac.py 2>&1 | tail -1 | python -c "import sys,json; j=json.loads(sys.stdin.read()); print({k:v for k,v in j['ms'].items() if 'jac' in k or 'lu' in k})"
