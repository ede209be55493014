# policy_ab.sh <variants...>: policy-kernel parity tests + tools/time_policy.py for build/variants/<name>.so, one box
mkdir -p gpurun_out
for v in "$@"; do
  echo "== $v"
  PUPPER_ENV_LIB=$PWD/build/variants/$v.so python -m pytest tests -m gpu -x -q -k "policy" 2>&1 | tail -2
  PUPPER_ENV_LIB=$PWD/build/variants/$v.so python tools/time_policy.py 8192 2>&1 | grep -v "^$" | tail -8
done
