"""CPU oracle for the CTAN hot path -- test infrastructure, not product code.
See oracle/README.md; the product package never imports anything from here."""
