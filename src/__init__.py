"""Drop-in alias of the reference's ``src`` package (optimiser, penalizers, result records)."""
