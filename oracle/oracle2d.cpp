// placeholder until the 2-D restatement lands
