# A/B of prebuilt library variants on one box: variants/lib_<name>.so are copied over the in-tree library in turn
set -u
for v in "$@"; do
  cp variants/lib_$v.so hvqm4_b200/libhvqm4_b200.so
  for rep in 1 2; do
    echo "== $v dense:     $(timeout 120 python tools/profile_recon.py 1024 3 0 2>&1 | tail -1)"
    echo "== $v realistic: $(timeout 120 python tools/profile_recon.py 1024 3 1 2>&1 | tail -1)"
  done
  echo "== $v sdk dense:     $(timeout 120 python tools/profile_sdk.py 0 2>&1 | tail -1)"
  echo "== $v sdk realistic: $(timeout 120 python tools/profile_sdk.py 1 2>&1 | tail -1)"
  echo "== $v parity: $(timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k 'golden or oracle_port' 2>&1 | tail -1)"
done
