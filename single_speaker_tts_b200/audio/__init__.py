"""Drop-in replacements for the reference's ``audio`` package functions on the hot path."""
