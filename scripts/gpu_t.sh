#!/bin/bash
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -1
fails=0
for i in $(seq 1 25); do
  python __graft_entry__.py --smoke > /tmp/smoke_$i.log 2>&1 || { fails=$((fails+1)); echo "run $i FAILED"; tail -12 /tmp/smoke_$i.log; }
done
echo "smoke failures: $fails / 25"
