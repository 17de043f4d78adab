"""Drop-in counterparts of the reference's ``model_zoo`` package (same module and class names)."""
