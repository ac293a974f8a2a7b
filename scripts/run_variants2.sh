bash scripts/run_variants.sh
echo "--- pool 16"
VARARGS="--pool 16" bash scripts/run_variants.sh
