#!/bin/bash
python -m pytest tests/test_second_order.py -m gpu -q 2>&1 | grep -E "passed|failed|^FAILED|Error|assert" | head -40
