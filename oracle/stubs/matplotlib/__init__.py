"""Import-only stand-in for matplotlib (absent in the image); the reference imports it at
module scope but the oracle never plots. Test infrastructure only."""
