#!/bin/bash
python -m pytest tests/test_second_order.py -m gpu -q 2>&1 | grep -E "passed|failed|^FAILED|Error" | head
python -m pytest tests/test_mmaml.py -m gpu -q -s 2>&1 | grep -E "passed|failed|^FAILED|mmaml/|mmaml2/" | head
