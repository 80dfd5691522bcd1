"""`python -m 2fast2q_b200 -c ...` == `2fast2q -c ...` (fast2q/__main__.py:1-3 of the reference)"""
from .fast2q import main

if __name__ == "__main__":
    main()
