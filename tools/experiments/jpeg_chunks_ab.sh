#!/bin/bash
python -m pytest tests/test_gpu_jpeg.py -m gpu -x -q 2>&1 | tail -3
for c in 1 2 4 8; do
  SPE_JPEG_CHUNKS=$c python bench.py --steps 20 --warmup 5 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); j=d['image_set']['from_jpeg_files']; print('chunks', $c, j.get('images_per_s'), j.get('seconds'), j.get('poses_solved_this_rank'), j.get('error'), 'frames:', round(d['image_set']['images_per_s']))"
done
