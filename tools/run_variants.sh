for v in nt256c2 nt128 nt128c5 nt192c3; do
  echo "== $v"
  BPPGPU_LIB=bpp_b200/variants/$v.so python tools/device_time.py config2
  BPPGPU_LIB=bpp_b200/variants/$v.so python tools/device_time.py config3 4000
done
