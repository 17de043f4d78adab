"""Drop-in counterparts of the hot-path pieces of the reference's ``dataloaders`` package."""
