use() { cp $1/libhifigan_b200.so $1/libhifigan_b200.srchash hifi-gan_b200/; cp $1/hg_resblock_pair.cu hifi-gan_b200/csrc/; }
t() { timeout 100 python tests/gpu_bringup.py pairs 64 1024 2>&1 | grep '"stage": "pair"' | grep '"d": 1,' | grep -v supported | python -c 'import json,sys; print(" ".join("C%d/k%d:%.3f" % (d["c"], d["k"], d["ms"]) for d in map(json.loads, sys.stdin)))'; }
for r in 1 2; do
  use ab_prev; echo "prev $(t)"
  use ab_new; echo "new  $(t)"
done
timeout 100 python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | tail -1
