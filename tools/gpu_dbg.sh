python -m tools.dbg_bounded_zc 2>&1 | tail -9
echo "--- with lam"; python -m tools.dbg_bounded_zc lam 2>&1 | tail -9
echo "--- with lam, MMPC_ZC_BIG=0"; MMPC_ZC_BIG=0 python -m tools.dbg_bounded_zc lam 2>&1 | tail -4
