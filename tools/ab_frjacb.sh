#!/bin/bash
# A/B of the reacting boundary-Jacobian kernel's register cap (one gpurun call)
for mb in 4 6 8; do PCFD_FRJACB_MINB=$mb timeout 200 python tools/time_frjac.py 2>&1 | tail -1; done
