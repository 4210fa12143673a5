"""Test-infrastructure shim (oracle only): lets the untouched reference modules import without matplotlib."""
