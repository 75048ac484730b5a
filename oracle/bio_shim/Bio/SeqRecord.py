class SeqRecord(object):
    def __init__(self, seq, id="<unknown id>", name="<unknown name>",
                 description="<unknown description>"):
        self.seq = seq
        self.id = id
        self.name = name
        self.description = description

    def __len__(self):
        return len(self.seq)
