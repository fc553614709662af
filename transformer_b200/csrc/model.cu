// placeholder -- replaced by the model orchestration
