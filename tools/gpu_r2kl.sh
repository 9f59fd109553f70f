#!/bin/bash
python tools/dbg_klement.py 2>&1 | tail -2 | cut -c1-400
timeout 600 python -m pytest tests -m gpu -q -k "klement or broyden or limited_memory" 2>&1 | tail -4
