python -m pytest tests/test_modules.py -m gpu -q -x 2>&1 | tail -3
echo "== two lanes"; python bench.py 2>/dev/null | cut -c1-200
echo "== one lane"; SVX_ENCODER_ONE_LANE=1 python bench.py 2>/dev/null | cut -c1-200
