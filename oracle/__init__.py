"""TEST INFRASTRUCTURE: CPU checkers of the likelihood path (C restatement, compiled reference, sampler port).
Nothing under nestfit_b200/ imports this package."""
