#!/bin/bash
# several table-diff runs: ragged small lengths, a full L=100 batch, L=200, L=400
python scripts/diff_fill.py 1 2>&1 | tail -8
python scripts/diff_fill.py 2 $(python -c "import random; random.seed(5); print(','.join(str(random.randint(1,140)) for _ in range(600)))") $1 2>&1 | tail -8
python scripts/diff_fill.py 3 $(python -c "print(','.join(['100']*1024))") $1 2>&1 | tail -6
python scripts/diff_fill.py 4 $(python -c "print(','.join(['200']*300+['199','201','197']))") $1 2>&1 | tail -6
python scripts/diff_fill.py 5 $(python -c "print(','.join(['400']*300+['399','398','397']))") $1 2>&1 | tail -6
