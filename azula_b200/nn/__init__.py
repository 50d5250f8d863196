r"""Neural-network helpers on the generation path."""
