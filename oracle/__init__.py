"""TEST INFRASTRUCTURE ONLY — see oracle/ref_engine.py header.  Never imported by xotorch_b200/."""
