python -m pytest tests/test_zsl.py -m gpu -x -q 2>&1 | tail -3
for v in default zepi8 zepi4s5 default zepi8; do if [ $v = default ]; then python scripts/dev_timing_zsl_quick.py 2>&1 | tail -1; else MRE_B200_LIB=$PWD/multimodal-relation-extrapolation_b200/build/variants/libmre_$v.so python scripts/dev_timing_zsl_quick.py 2>&1 | tail -1; fi; done
