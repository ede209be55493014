for v in "$@"; do
  echo "== $v"
  PUPPER_ENV_LIB=$PWD/build/variants/$v.so timeout 150 python -m pytest tests -m gpu -x -q -k "policy_kernel" 2>&1 | grep -E "passed|failed|AssertionError" | head -3
done
