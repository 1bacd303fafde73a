#!/bin/bash
# End-to-end throughput of qr_step_host against the number of pipeline chunks (QR_HOST_CHUNKS override): bash tools/e2e_chunks.sh
cd /root/repo
for c in 4 8 12 16 24 32; do
  for rep in 1 2; do
    QR_HOST_CHUNKS=$c timeout 200 python bench.py --no-cpu --no-extra --steps 3 --warmup 3 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); e=d['e2e']; print('chunks $c', 'e2e %.1f M  ceiling %.1f M  ratio %.3f'%(e['value']/1e6, e['copy_ceiling']/1e6, e['value']/e['copy_ceiling']))"
  done
done
