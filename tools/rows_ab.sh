# A/B of the row-block K1: rows per CTA x CTAs per SM (PTK_KCOST_CFG = 0: 6, 5, 4)
for rows in ${ROWS:-8 16}; do for cfg in ${CFGS:-0 5 4}; do
  echo "rows=$rows cfg=$cfg"; PTK_KCOST_ROWS=$rows PTK_KCOST_CFG=$cfg python tools/kcost_ab.py child
done; done
