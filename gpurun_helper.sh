set +x
cp slam-localization_b200/csrc/libslb.so /tmp/libslb_orig.so
make -C slam-localization_b200/csrc timing CALL=1 EXTRA=-DSLB_EXP_NOK > /dev/null 2>&1
cp slam-localization_b200/csrc/libslb_timing.so slam-localization_b200/csrc/libslb.so
timeout 120 python profiles/chol_timing.py ukf slam-localization_b200/csrc/libslb.so 2>&1 | tail -15
cp /tmp/libslb_orig.so slam-localization_b200/csrc/libslb.so
