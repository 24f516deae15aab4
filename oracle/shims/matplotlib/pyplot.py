"""TEST INFRASTRUCTURE ONLY -- empty stand-in."""
