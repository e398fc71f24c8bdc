# tuning lottery: the cfg3 kernel object compiled with different flags, same sources (NDNLM_LIB selects the library)
for v in "" _v1 _v2 _v3 _v4 _v5; do
  echo "== libndnlm$v.so"
  NDNLM_LIB=$PWD/nd_b200/libndnlm$v.so python tools/dev_multi.py 0 --shape 296,4096,32,4 --steps 3 2>&1 | tail -1
done
