"""Drop-in module path of the reference's scheduler package (its test_*.py do `from scheduler... import ...`)."""
