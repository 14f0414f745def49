"""Audio-side replacements for the reference's ``datasets`` package (feature recipe, statistics)."""
