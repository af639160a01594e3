for rows in 0 4 8 16; do for cfg in 0 15 14; do
  if [ $rows = 0 ] && [ $cfg != 0 ]; then continue; fi
  echo "rows=$rows cfg=$cfg"; PTK_KCOST_ROWS=$rows PTK_KCOST_CFG=$cfg python tools/kcost_ab.py child
done; done
