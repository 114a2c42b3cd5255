"""TEST INFRASTRUCTURE ONLY: CPU oracle for the vitron_b200 parity tests (never imported by the product)."""
