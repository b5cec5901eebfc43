class GradientState:
    end_of_dataloader = False
    remainder = -1
