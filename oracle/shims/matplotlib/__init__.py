"""TEST INFRASTRUCTURE ONLY -- empty stand-in: matplotlib is not installed; the reference imports pyplot at module level."""
