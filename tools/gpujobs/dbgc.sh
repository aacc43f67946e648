#!/bin/bash
cd /root/repo
python -m pytest tests/test_gpu_contact_debug.py tests/test_gpu_parity.py -q 2>&1 | tail -6
python tests/rollout_report.py --help > /dev/null 2>&1; echo "rollout_report import rc=$?"
