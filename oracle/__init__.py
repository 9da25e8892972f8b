"""CPU float64 oracle of the maze step path — TEST INFRASTRUCTURE, never imported by the product."""
