"""TEST INFRASTRUCTURE ONLY — CPU oracle of the pmoe_b200 hot path (see oracle/functional.py).
Nothing under pmoe_b200/ may import this package."""
