class AbstractRenormalization:
    pass


class SumOfOperators:
    pass
