"""matplotlib is not installed; examples/stateFollowingHO.py only imports pyplot."""
