"""Host-side mirror of the reference's `cell_type_annotation` package (same module and class names,
same call signatures) whose hot path runs in libribca_b200.so."""
