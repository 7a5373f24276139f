#!/bin/bash
# retry a gpurun command until the pod has a free slot (transient answers are not charged)
for i in $(seq 1 40); do
  out=$(gpurun --timeout 1500 -- "$1" 2>&1)
  echo "$out" | tail -25
  if ! echo "$out" | grep -q "status=transient"; then exit 0; fi
  sleep 45
done
