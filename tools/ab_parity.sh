#!/bin/bash
# tools/ab_parity.sh <variant.so> <tag>: step-path timing + the Contract-X parity figures of one build of the library (GPU box)
lib=$1; tag=$2
export TVC_B200_LIB=$PWD/$lib
python tools/ab_step.py $tag 2>&1 | grep AB
python -m pytest tests/test_gpu_parity.py -m gpu -q -k "contract_x or full_size or random_autoreset" 2>&1 | grep "parity\]\|passed\|failed" | python -c "
import sys, json
for l in sys.stdin:
    if '[parity]' in l:
        name, js = l.split(': ', 1); d = json.loads(js)
        if 'contact_max' in d: print('$tag', name.split()[-1], 'free', '%.2e' % d.get('free_flight_max', 0), 'q99', '%.2e' % d.get('contact_q99', 0), 'max', '%.2e' % d['contact_max'], 'flags', d.get('flag_mismatches'))
    else: print('$tag', l.strip())
"
